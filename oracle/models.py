"""ORACLE (test infrastructure, never imported by the product path).

Torch-CPU fp32 definitions of the four third-party graphs on the identity hot path
(SURVEY.md App. A.5): SCRFD-10G-bnkps, SCRFD-2.5G-bnkps, ArcFace iResNet-50 / iResNet-100.
The reference runs them as ONNX files through ONNX Runtime / TensorRT
(person_capture/face_embedder.py:1102-1107 for SCRFD, :891-915 and :1369 for ArcFace); the
ONNX files themselves are downloaded at run time (face_embedder.py:55-83) and are absent
here, so this file *defines* the architectures from the upstream InsightFace configs and
both sides (oracle and CUDA path) load the same exported weight file.

parity unpinned for the graphs: the reference ships no tests or golden vectors for them
(SURVEY.md F2); torch-CPU fp32 stands in for ONNX Runtime CPU (not installed, F4).  The graphs ARE exported to ONNX
(oracle/onnx_export.py writes the protobuf by hand -- the `onnx` package is absent too) and executed by an independent engine,
cv2.dnn: tests/test_cpu_onnx_export.py holds these torch executors to it for all four networks.

Weight file format ("folded", one .npz per model): every conv `name` has
  name.w      float16 [Cout, Cin, kh, kw]   (values exactly representable in fp16)
  name.scale  float32 [Cout]                (BN gamma/sqrt(var+eps), or 1)
  name.bias   float32 [Cout]                (BN beta-mean*scale, or conv bias)
  name.slope  float32 [Cout]                (PReLU only)
and every standalone BN `name` has name.scale / name.bias.  The forward below in "folded"
mode computes  act(scale * conv(x, w) + bias [+ residual])  in fp32, which is exactly the
arithmetic the CUDA epilogues perform.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

BN_EPS = 1e-5

SCRFD_CFG = {
    # name: stem (c1, c2, c3), stage blocks, stage planes, fpn out, head convs, head channels
    "scrfd_10g_bnkps": dict(stem=(28, 28, 56), blocks=(3, 4, 2, 3), planes=(56, 88, 88, 224), fpn=56, stacked=3, feat=80),
    "scrfd_2.5g_bnkps": dict(stem=(12, 12, 24), blocks=(3, 5, 3, 2), planes=(24, 48, 48, 80), fpn=24, stacked=2, feat=64),
}
IRESNET_CFG = {
    "arcface_r50": (3, 4, 14, 3),
    "arcface_r100": (3, 13, 30, 3),
}


class ConvBN(nn.Module):
    """conv (no bias) + BN, optional ReLU.  Exported as one folded conv."""

    def __init__(self, cin, cout, k, stride=1, act="relu", bn=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, k // 2, bias=not bn)
        self.bn = nn.BatchNorm2d(cout, eps=BN_EPS) if bn else None
        self.act = act

    def forward(self, x, residual=None):
        y = self.conv(x)
        if self.bn is not None:
            y = self.bn(y)
        if residual is not None:
            y = y + residual
        if self.act == "relu":
            y = F.relu(y)
        return y

    def folded(self, prefix: str, out: Dict[str, np.ndarray]):
        w = self.conv.weight.detach()
        out[prefix + ".w"] = w.to(torch.float16).numpy()
        if self.bn is not None:
            s = self.bn.weight.detach() / torch.sqrt(self.bn.running_var + self.bn.eps)
            b = self.bn.bias.detach() - self.bn.running_mean * s
        else:
            s = torch.ones(w.shape[0])
            b = self.conv.bias.detach()
        out[prefix + ".scale"] = s.float().numpy()
        out[prefix + ".bias"] = b.float().numpy()


class BasicBlockV1e(nn.Module):
    def __init__(self, cin, planes, stride):
        super().__init__()
        self.conv1 = ConvBN(cin, planes, 3, stride, "relu")
        self.conv2 = ConvBN(planes, planes, 3, 1, "relu")  # relu applied after residual add
        self.down = None
        self.stride = stride
        if stride != 1 or cin != planes:
            self.down = ConvBN(cin, planes, 1, 1, "none")

    def forward(self, x):
        idt = x
        if self.down is not None:
            if self.stride != 1:
                idt = F.avg_pool2d(x, self.stride, self.stride, ceil_mode=True, count_include_pad=False)
            idt = self.down(idt)
        y = self.conv1(x)
        return self.conv2(y, residual=idt)


class SCRFD(nn.Module):
    """ResNetV1e (deep stem, avg-down) + PAFPN + shared BN head with 2 anchors and 5 kps."""

    strides = (8, 16, 32)

    def __init__(self, name: str):
        super().__init__()
        cfg = SCRFD_CFG[name]
        self.name = name
        c1, c2, c3 = cfg["stem"]
        self.stem1 = ConvBN(3, c1, 3, 2)
        self.stem2 = ConvBN(c1, c2, 3, 1)
        self.stem3 = ConvBN(c2, c3, 3, 1)
        layers = []
        cin = c3
        self.stage_names = []
        for si, (nb, planes) in enumerate(zip(cfg["blocks"], cfg["planes"])):
            blocks = []
            for bi in range(nb):
                stride = 2 if (bi == 0 and si > 0) else 1
                blocks.append(BasicBlockV1e(cin, planes, stride))
                cin = planes
            layers.append(nn.ModuleList(blocks))
        self.stages = nn.ModuleList(layers)
        fo = cfg["fpn"]
        pin = cfg["planes"][1:]
        self.lateral = nn.ModuleList([ConvBN(c, fo, 1, 1, "none", bn=False) for c in pin])
        self.fpn = nn.ModuleList([ConvBN(fo, fo, 3, 1, "none", bn=False) for _ in pin])
        self.down = nn.ModuleList([ConvBN(fo, fo, 3, 2, "none", bn=False) for _ in range(2)])
        self.pafpn = nn.ModuleList([ConvBN(fo, fo, 3, 1, "none", bn=False) for _ in range(2)])
        fc = cfg["feat"]
        tower = []
        for i in range(cfg["stacked"]):
            tower.append(ConvBN(fo if i == 0 else fc, fc, 3, 1, "relu"))
        self.tower = nn.ModuleList(tower)
        # fused output conv: channels = [cls a0,a1 | reg a0(4),a1(4) | kps a0(10),a1(10)] = 30
        self.out = ConvBN(fc, 30, 3, 1, "none", bn=False)
        self.reg_scale = nn.Parameter(torch.ones(3))
        nn.init.constant_(self.out.conv.bias[:2], -4.0)

    def features(self, x):
        x = self.stem3(self.stem2(self.stem1(x)))
        x = F.max_pool2d(x, 3, 2, 1)
        outs = []
        for st in self.stages:
            for b in st:
                x = b(x)
            outs.append(x)
        c = outs[1:]
        lat = [l(t) for l, t in zip(self.lateral, c)]
        for i in (2, 1):
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], scale_factor=2, mode="nearest")
        inter = [f(t) for f, t in zip(self.fpn, lat)]
        for i in range(2):
            inter[i + 1] = inter[i + 1] + self.down[i](inter[i])
        return [inter[0], self.pafpn[0](inter[1]), self.pafpn[1](inter[2])]

    def head_raw(self, x):
        """Per level raw maps [N,30,h,w] (cls logits, reg*scale, kps)."""
        res = []
        for li, f in enumerate(self.features(x)):
            t = f
            for m in self.tower:
                t = m(t)
            o = self.out(t)
            o = torch.cat([o[:, :2], o[:, 2:10] * self.reg_scale[li], o[:, 10:]], dim=1)
            res.append(o)
        return res

    def forward(self, x):
        """The 9 ONNX outputs for a single image: score_8,16,32, bbox_8,16,32, kps_8,16,32.

        Rows are ordered (y, x, anchor); score is post-sigmoid.
        """
        raws = self.head_raw(x)
        sc, bb, kp = [], [], []
        for o in raws:
            o = o[0].permute(1, 2, 0)  # h, w, 30
            h, w, _ = o.shape
            sc.append(torch.sigmoid(o[..., 0:2]).reshape(h * w * 2, 1))
            bb.append(o[..., 2:10].reshape(h * w * 2, 4))
            kp.append(o[..., 10:30].reshape(h * w * 2, 10))
        return tuple(sc + bb + kp)

    def export_folded(self) -> Dict[str, np.ndarray]:
        out: Dict[str, np.ndarray] = {}
        self.stem1.folded("stem1", out)
        self.stem2.folded("stem2", out)
        self.stem3.folded("stem3", out)
        for si, st in enumerate(self.stages):
            for bi, b in enumerate(st):
                p = f"s{si}.b{bi}"
                b.conv1.folded(p + ".conv1", out)
                b.conv2.folded(p + ".conv2", out)
                if b.down is not None:
                    b.down.folded(p + ".down", out)
        for i in range(3):
            self.lateral[i].folded(f"lateral{i}", out)
            self.fpn[i].folded(f"fpn{i}", out)
        for i in range(2):
            self.down[i].folded(f"down{i}", out)
            self.pafpn[i].folded(f"pafpn{i}", out)
        for i, m in enumerate(self.tower):
            m.folded(f"tower{i}", out)
        self.out.folded("out", out)
        out["reg_scale"] = self.reg_scale.detach().float().numpy()
        return out


# --------------------------------------------------------------------------------------
# Folded-parameter executors (what parity is measured against).
# --------------------------------------------------------------------------------------

def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


class FoldedNet:
    """Runs a folded .npz on torch-CPU fp32: y = act(scale*conv(x,w)+bias [+res])."""

    def __init__(self, params: Dict[str, np.ndarray]):
        self.p = {k: _t(np.asarray(v)).float() for k, v in params.items()}

    def conv(self, x, name, stride=1, act="none", residual=None):
        w = self.p[name + ".w"]
        k = w.shape[-1]
        y = F.conv2d(x, w, None, stride, k // 2)
        y = y * self.p[name + ".scale"].view(1, -1, 1, 1) + self.p[name + ".bias"].view(1, -1, 1, 1)
        if residual is not None:
            y = y + residual
        if act == "relu":
            y = F.relu(y)
        elif act == "prelu":
            y = torch.where(y >= 0, y, y * self.p[name + ".slope"].view(1, -1, 1, 1))
        return y

    def affine(self, x, name):
        return x * self.p[name + ".scale"].view(1, -1, 1, 1) + self.p[name + ".bias"].view(1, -1, 1, 1)


class FoldedSCRFD(FoldedNet):
    strides = (8, 16, 32)

    def __init__(self, name: str, params):
        super().__init__(params)
        self.cfg = SCRFD_CFG[name]
        self.name = name

    @torch.no_grad()
    def head_raw(self, x: torch.Tensor) -> List[torch.Tensor]:
        cfg = self.cfg
        x = self.conv(x, "stem1", 2, "relu")
        x = self.conv(x, "stem2", 1, "relu")
        x = self.conv(x, "stem3", 1, "relu")
        x = F.max_pool2d(x, 3, 2, 1)
        outs = []
        cin = cfg["stem"][2]
        for si, (nb, planes) in enumerate(zip(cfg["blocks"], cfg["planes"])):
            for bi in range(nb):
                stride = 2 if (bi == 0 and si > 0) else 1
                p = f"s{si}.b{bi}"
                idt = x
                if stride != 1 or cin != planes:
                    if stride != 1:
                        idt = F.avg_pool2d(x, 2, 2)
                    idt = self.conv(idt, p + ".down", 1, "none")
                y = self.conv(x, p + ".conv1", stride, "relu")
                x = self.conv(y, p + ".conv2", 1, "relu", residual=idt)
                cin = planes
            outs.append(x)
        c = outs[1:]
        lat = [self.conv(t, f"lateral{i}") for i, t in enumerate(c)]
        for i in (2, 1):
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], scale_factor=2, mode="nearest")
        inter = [self.conv(t, f"fpn{i}") for i, t in enumerate(lat)]
        for i in range(2):
            inter[i + 1] = inter[i + 1] + self.conv(inter[i], f"down{i}", 2)
        feats = [inter[0], self.conv(inter[1], "pafpn0"), self.conv(inter[2], "pafpn1")]
        res = []
        for li, f in enumerate(feats):
            t = f
            for i in range(cfg["stacked"]):
                t = self.conv(t, f"tower{i}", 1, "relu")
            res.append(self.conv(t, "out"))
        return res

    @torch.no_grad()
    def run(self, blob: np.ndarray):
        """blob float32 [1,3,S,S] -> the 9 ONNX outputs as numpy."""
        raws = self.head_raw(_t(blob).float())
        rs = self.p["reg_scale"]
        sc, bb, kp = [], [], []
        for li, o in enumerate(raws):
            o = o[0].permute(1, 2, 0)
            h, w, _ = o.shape
            sc.append(torch.sigmoid(o[..., 0:2]).reshape(h * w * 2, 1).numpy())
            bb.append((o[..., 2:10] * rs[li]).reshape(h * w * 2, 4).numpy())
            kp.append(o[..., 10:30].reshape(h * w * 2, 10).numpy())
        return sc + bb + kp


class FoldedIResNet(FoldedNet):
    def __init__(self, name: str, params):
        super().__init__(params)
        self.blocks = IRESNET_CFG[name]
        self.name = name

    @torch.no_grad()
    def run(self, x: np.ndarray, batch: int = 16) -> np.ndarray:
        """x float32 [n,3,112,112] ((RGB-127.5)/127.5) -> raw embeddings float32 [n,512]."""
        outs = []
        for i in range(0, x.shape[0], batch):
            outs.append(self._fwd(_t(x[i:i + batch]).float()).numpy())
        return np.concatenate(outs, 0) if outs else np.zeros((0, 512), np.float32)

    def _fwd(self, x):
        x = self.conv(x, "stem", 1, "prelu")
        cin = 64
        for si, nb in enumerate(self.blocks):
            planes = 64 << si
            for bi in range(nb):
                p = f"s{si}.b{bi}"
                stride = 2 if bi == 0 else 1
                if bi == 0:
                    idt = self.conv(x, p + ".down", stride, "none")
                else:
                    idt = x
                y = self.affine(x, p + ".bn1")
                y = self.conv(y, p + ".conv1", 1, "prelu")
                x = self.conv(y, p + ".conv2", stride, "none", residual=idt)
                cin = planes
        x = self.affine(x, "bn2")
        n = x.shape[0]
        # the FC weight is stored for an (h, w, c) flatten so both sides share one layout
        x = x.permute(0, 2, 3, 1).reshape(n, -1)
        y = x @ self.p["fc.w"].t()
        return y * self.p["fc.scale"] + self.p["fc.bias"]


# --------------------------------------------------------------------------------------
# ArcFace: seeded random weights (person_capture_b200/weights.py) + data-dependent affine terms.
# --------------------------------------------------------------------------------------

def calibrate_arcface_affine(name: str, W: Dict[str, np.ndarray], calib: np.ndarray, seed: int = 777) -> Dict[str, np.ndarray]:
    """Per-layer scale/bias/slope for the random iResNet `W`, from calibration statistics.

    Every folded scale/bias is chosen so that the layer output over `calib` (float32
    [n,3,112,112]) has zero mean / unit variance per channel, the residual branch is damped
    (0.35) and the final 512-d feature is centred and whitened per dimension.  That keeps a
    100-layer random net well conditioned in fp16 and makes distinct textures decorrelate
    (cosine ~ 0) while near-identical chips stay close -- the property the fd thresholds
    (gui_app.py:561-563) rely on.  Output contains only the small arrays (committed).
    """
    rng = np.random.default_rng(seed)
    P: Dict[str, np.ndarray] = {}
    x = _t(calib).float()

    def stats(y):
        return y.mean(dim=(0, 2, 3)), y.var(dim=(0, 2, 3), unbiased=False)

    def add_conv(pname, x_in, stride, damp=1.0, slope=False):
        w = _t(W[pname + ".w"]).float()
        k = w.shape[-1]
        y = F.conv2d(x_in, w, None, stride, k // 2)
        m, v = stats(y)
        s = damp / torch.sqrt(v + 1e-5)
        b = -m * s
        P[pname + ".scale"] = s.numpy().astype(np.float32)
        P[pname + ".bias"] = b.numpy().astype(np.float32)
        y = y * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
        if slope:
            sl = rng.uniform(0.1, 0.3, size=w.shape[0]).astype(np.float32)
            P[pname + ".slope"] = sl
            y = torch.where(y >= 0, y, y * _t(sl).view(1, -1, 1, 1))
        return y

    with torch.no_grad():
        x = add_conv("stem", x, 1, slope=True)
        for si, nb in enumerate(IRESNET_CFG[name]):
            for bi in range(nb):
                p = f"s{si}.b{bi}"
                stride = 2 if bi == 0 else 1
                idt = add_conv(p + ".down", x, stride) if bi == 0 else x
                m, v = stats(x)
                s = 1.0 / torch.sqrt(v + 1e-5)
                P[p + ".bn1.scale"] = s.numpy().astype(np.float32)
                P[p + ".bn1.bias"] = (-m * s).numpy().astype(np.float32)
                y = x * s.view(1, -1, 1, 1) + (-m * s).view(1, -1, 1, 1)
                y = add_conv(p + ".conv1", y, 1, slope=True)
                y = add_conv(p + ".conv2", y, stride, damp=0.35)
                x = y + idt
        m, v = stats(x)
        s = 1.0 / torch.sqrt(v + 1e-5)
        P["bn2.scale"] = s.numpy().astype(np.float32)
        P["bn2.bias"] = (-m * s).numpy().astype(np.float32)
        x = x * s.view(1, -1, 1, 1) + (-m * s).view(1, -1, 1, 1)
        n = x.shape[0]
        flat = x.permute(0, 2, 3, 1).reshape(n, -1)
        y = flat @ _t(W["fc.w"]).float().t()
        m = y.mean(0)
        v = y.var(0, unbiased=False)
        s = 1.0 / torch.sqrt(v + 1e-5)
        P["fc.scale"] = s.numpy().astype(np.float32)
        P["fc.bias"] = (-m * s).numpy().astype(np.float32)
    return P
