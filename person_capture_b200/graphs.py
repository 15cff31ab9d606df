"""Op tables of the four graphs on the path, in the form pcb_model_load() consumes.

The reference executes downloaded ONNX files (person_capture/face_embedder.py:55-83) through
ONNX Runtime; here the graphs are defined from the upstream architectures
as flat op lists over numbered tensors, and FLOPs are counted from these tables
(`graph_macs`), never from the model name.  Weight names match weights.load_params().
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import numpy as np

from . import _lib as L

SCRFD_CFG = {
    "scrfd_10g_bnkps": dict(stem=(28, 28, 56), blocks=(3, 4, 2, 3), planes=(56, 88, 88, 224), fpn=56, stacked=3, feat=80),
    "scrfd_2.5g_bnkps": dict(stem=(12, 12, 24), blocks=(3, 5, 3, 2), planes=(24, 48, 48, 80), fpn=24, stacked=2, feat=64),
}
IRESNET_BLOCKS = {"arcface_r50": (3, 4, 14, 3), "arcface_r100": (3, 13, 30, 3)}


class Graph:
    def __init__(self, name: str):
        self.name = name
        self.ops: List[dict] = []
        self.n_tensors = 1
        self.outputs: List[int] = []
        self.tensor_names: Dict[str, int] = {"input": 0}
        self.scale_of: Dict[int, float] = {0: 1.0}   # spatial size of tensor relative to tensor 0 (for MAC counts)
        self.chan_of: Dict[int, int] = {0: 27}

    def new_tensor(self, label: str, rel: float, ch: int) -> int:
        t = self.n_tensors
        self.n_tensors += 1
        self.tensor_names[label] = t
        self.scale_of[t] = rel
        self.chan_of[t] = ch
        return t

    def conv(self, label, x, cin, cout, k, stride, act, wname, residual=-1, out_f32=False, affine2=None):
        """affine2: name of a BatchNorm whose scale/bias are applied to this conv's (activated) output and
        written as a second fp16 tensor; returns (out, out2) in that case."""
        stem = x == 0
        rel = self.scale_of[x] / (1 if stem else stride)
        out = self.new_tensor(label, rel, cout)
        out2 = self.new_tensor(label + "+" + affine2, rel, cout) if affine2 else -1
        self.ops.append(dict(kind=L.OP_CONV, in0=x, in1=residual, out=out, cin=cin, cout=cout, k=k, stride=stride,
                             act=act, wname=wname, out2=out2, flags=L.OPF_OUT_F32 if out_f32 else 0, w2name=affine2))
        return (out, out2) if affine2 else out

    def simple(self, kind, label, x, x2=-1, rel_mul=1.0, wname=None, cout=None):
        out = self.new_tensor(label, self.scale_of[x] * rel_mul, cout or self.chan_of[x])
        self.ops.append(dict(kind=kind, in0=x, in1=x2, out=out, cin=self.chan_of[x], cout=cout or self.chan_of[x], k=0,
                             stride=0, act=0, wname=wname, out2=-1, flags=0))
        return out


def scrfd_graph(name: str) -> Graph:
    cfg = SCRFD_CFG[name]
    g = Graph(name)
    c1, c2, c3 = cfg["stem"]
    x = g.conv("stem1", 0, 3, c1, 3, 2, L.ACT_RELU, "stem1")
    x = g.conv("stem2", x, c1, c2, 3, 1, L.ACT_RELU, "stem2")
    x = g.conv("stem3", x, c2, c3, 3, 1, L.ACT_RELU, "stem3")
    x = g.simple(L.OP_MAXPOOL3S2, "pool", x, rel_mul=0.5)
    cin = c3
    stage_out = []
    for si, (nb, planes) in enumerate(zip(cfg["blocks"], cfg["planes"])):
        for bi in range(nb):
            stride = 2 if (bi == 0 and si > 0) else 1
            p = f"s{si}.b{bi}"
            idt = x
            if stride != 1 or cin != planes:
                if stride != 1:
                    idt = g.simple(L.OP_AVGPOOL2, p + ".avg", x, rel_mul=0.5)
                idt = g.conv(p + ".down", idt, cin, planes, 1, 1, L.ACT_NONE, p + ".down")
            y = g.conv(p + ".conv1", x, cin, planes, 3, stride, L.ACT_RELU, p + ".conv1")
            x = g.conv(p + ".conv2", y, planes, planes, 3, 1, L.ACT_RELU, p + ".conv2", residual=idt)
            cin = planes
        stage_out.append(x)
    fo = cfg["fpn"]
    c = stage_out[1:]
    lat = [g.conv(f"lateral{i}", c[i], cfg["planes"][1 + i], fo, 1, 1, L.ACT_NONE, f"lateral{i}") for i in range(3)]
    lat[1] = g.simple(L.OP_UPSAMPLE_ADD, "td1", lat[1], lat[2])
    lat[0] = g.simple(L.OP_UPSAMPLE_ADD, "td0", lat[0], lat[1])
    inter = [g.conv(f"fpn{i}", lat[i], fo, fo, 3, 1, L.ACT_NONE, f"fpn{i}") for i in range(3)]
    # bottom-up: inter[i+1] += down_i(inter[i])  (residual of the stride-2 conv)
    inter[1] = g.conv("down0", inter[0], fo, fo, 3, 2, L.ACT_NONE, "down0", residual=inter[1])
    inter[2] = g.conv("down1", inter[1], fo, fo, 3, 2, L.ACT_NONE, "down1", residual=inter[2])
    feats = [inter[0], g.conv("pafpn0", inter[1], fo, fo, 3, 1, L.ACT_NONE, "pafpn0"),
             g.conv("pafpn1", inter[2], fo, fo, 3, 1, L.ACT_NONE, "pafpn1")]
    fc = cfg["feat"]
    for li, f in enumerate(feats):
        t = f
        for i in range(cfg["stacked"]):
            t = g.conv(f"l{li}.tower{i}", t, fo if i == 0 else fc, fc, 3, 1, L.ACT_RELU, f"tower{i}")
        # head maps stay fp32: fp16 would quantise 8..16-stride-unit distances to 2^-7 (0.25-0.5 px at stride 32)
        g.outputs.append(g.conv(f"l{li}.out", t, fc, 30, 3, 1, L.ACT_NONE, "out", out_f32=True))
    return g


def iresnet_graph(name: str) -> Graph:
    """IBasicBlock: bn1 -> conv1+bn+prelu -> conv2(stride)+bn, + shortcut.  The residual stream stays in fp32
    inside a stage (so fp16 rounding does not random-walk over 3..30 blocks), and every producer of the
    stream also emits bn1 of the *next* block as a second fp16 output (no separate BatchNorm pass)."""
    import os
    # measured on B200 (tools/fd_error.py): ArcFace-only |d fd| vs the fp32 oracle is 1.9e-4 max with an fp16
    # stream and 1.6e-4 with an fp32 one, both far inside the 1e-3 bar, while the fp32 stream halves the
    # throughput of the residual layers -> fp16 stream by default
    f32_stream = os.environ.get("PCB_F32_STREAM", "0") == "1"
    g = Graph(name)
    blocks = [(si, bi, nb) for si, nb in enumerate(IRESNET_BLOCKS[name]) for bi in range(nb)]
    x, y = g.conv("stem", 0, 3, 64, 3, 1, L.ACT_PRELU, "stem", affine2="s0.b0.bn1")   # x fp16 (feeds s0.b0.down)
    cin = 64
    for k, (si, bi, nb) in enumerate(blocks):
        planes = 64 << si
        p = f"s{si}.b{bi}"
        stride = 2 if bi == 0 else 1
        # identity path: stage-first blocks project the fp16 stage input; others read the fp32 stream
        idt = g.conv(p + ".down", x, cin, planes, 1, stride, L.ACT_NONE, p + ".down", out_f32=f32_stream) if bi == 0 else x
        y = g.conv(p + ".conv1", y, cin, planes, 3, 1, L.ACT_PRELU, p + ".conv1")
        last_in_stage = bi == nb - 1
        nxt = blocks[k + 1] if k + 1 < len(blocks) else None
        if nxt is not None:
            x, y = g.conv(p + ".conv2", y, planes, planes, 3, stride, L.ACT_NONE, p + ".conv2", residual=idt,
                          out_f32=f32_stream and not last_in_stage, affine2=f"s{nxt[0]}.b{nxt[1]}.bn1")
        else:
            x = g.conv(p + ".conv2", y, planes, planes, 3, stride, L.ACT_NONE, p + ".conv2", residual=idt)
        cin = planes
    flat = g.simple(L.OP_AFFINE_FLATTEN, "bn2_flat", x, wname="bn2", cout=512)
    g.ops.append(dict(kind=L.OP_FC, in0=flat, in1=-1, out=-1, cin=7 * 7 * 512, cout=512, k=1, stride=1, act=0, wname="fc",
                      out2=-1, flags=0))
    return g


def build_graph(name: str) -> Graph:
    return scrfd_graph(name) if name.startswith("scrfd") else iresnet_graph(name)


def graph_macs(g: Graph, in_h: int, in_w: int) -> int:
    """Dense multiply-accumulates of the conv/FC ops for one image whose patch tensor (tensor 0)
    is in_h x in_w (SCRFD: S/2 x S/2, ArcFace: 112 x 112)."""
    total = 0
    for op in g.ops:
        if op["kind"] == L.OP_CONV:
            rel = g.scale_of[op["out"]]
            total += int(round(in_h * rel)) * int(round(in_w * rel)) * op["cout"] * op["cin"] * op["k"] * op["k"]
        elif op["kind"] == L.OP_FC:
            total += op["cin"] * op["cout"]
    return total


def pack(g: Graph, params: Dict[str, np.ndarray]) -> Tuple[C.Array, bytes, C.Array, np.ndarray | None]:
    """-> (pcb_op array, blob bytes, outputs array, reg_scale or None)."""
    chunks: List[bytes] = []
    offset = 0
    cache: Dict[Tuple[str, str], int] = {}

    def put(wname: str, suffix: str, arr: np.ndarray) -> int:
        nonlocal offset
        key = (wname, suffix)
        if key in cache:
            return cache[key]
        b = np.ascontiguousarray(arr).tobytes()
        padn = (-len(b)) % 16
        cache[key] = offset
        chunks.append(b + b"\0" * padn)
        off = offset
        offset += len(b) + padn
        return off

    ops = (L.PcbOp * len(g.ops))()
    for i, op in enumerate(g.ops):
        o = ops[i]
        o.kind, o.in0, o.in1, o.out = op["kind"], op["in0"], op["in1"], op["out"]
        o.cin, o.cout, o.k, o.stride, o.act = op["cin"], op["cout"], op["k"], op["stride"], op["act"]
        o.w_off = o.scale_off = o.bias_off = o.slope_off = o.scale2_off = o.bias2_off = -1
        o.out2 = op.get("out2", -1)
        o.flags = op.get("flags", 0)
        wn = op["wname"]
        if op["kind"] in (L.OP_CONV, L.OP_FC):
            w = np.asarray(params[wn + ".w"], dtype=np.float16)
            exp = (op["cout"], op["cin"], op["k"], op["k"]) if op["kind"] == L.OP_CONV else (op["cout"], op["cin"])
            if tuple(w.shape) != exp:
                raise ValueError(f"{g.name}:{wn}: weight shape {w.shape} != {exp}")
            o.w_off = put(wn, "w", w)
            o.scale_off = put(wn, "scale", np.asarray(params[wn + ".scale"], np.float32))
            o.bias_off = put(wn, "bias", np.asarray(params[wn + ".bias"], np.float32))
            if op["act"] == L.ACT_PRELU:
                o.slope_off = put(wn, "slope", np.asarray(params[wn + ".slope"], np.float32))
            if op.get("w2name"):
                o.scale2_off = put(op["w2name"], "scale", np.asarray(params[op["w2name"] + ".scale"], np.float32))
                o.bias2_off = put(op["w2name"], "bias", np.asarray(params[op["w2name"] + ".bias"], np.float32))
        elif op["kind"] in (L.OP_AFFINE, L.OP_AFFINE_FLATTEN):
            o.scale_off = put(wn, "scale", np.asarray(params[wn + ".scale"], np.float32))
            o.bias_off = put(wn, "bias", np.asarray(params[wn + ".bias"], np.float32))
    outs = (C.c_int32 * max(1, len(g.outputs)))(*g.outputs)
    reg = np.asarray(params["reg_scale"], np.float32) if "reg_scale" in params else None
    return ops, b"".join(chunks), outs, reg
