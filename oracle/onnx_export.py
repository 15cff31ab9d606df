"""ORACLE (test infrastructure): writes the folded networks of oracle/models.py as ONNX files.

north_star asks for "identical random-init weights exported to ONNX" so that the reference's ONNX Runtime sessions
(person_capture/face_embedder.py:1102-1107 SCRFD, :891-915 ArcFace) can run the very graphs the CUDA path runs.  Neither
`onnx` nor `onnxruntime` exists offline, so the protobuf wire format is written by hand (the handful of message types an
inference graph needs: ModelProto, GraphProto, NodeProto, AttributeProto, TensorProto, ValueInfoProto; opset 13) and the result
is executed by an independent engine that IS here, OpenCV's `cv2.dnn.readNetFromONNX` (tests/test_cpu_onnx_export.py holds the
torch-CPU executors of oracle/models.py to it).

  export_scrfd(name, params, S, path, layout="insightface" | "raw")
      "insightface": the nine outputs InsightFace's SCRFD wrapper consumes (scores [A,1], boxes [A,4], landmarks [A,10] per
      stride, A = (S/stride)^2 * 2), fixed input [1,3,S,S]; "raw": the three [1,30,h,w] head maps
  export_iresnet(name, params, path, batch=1)
      input [batch,3,112,112] ((RGB - 127.5) / 127.5) -> embedding [batch,512]

Folding: y = scale * conv(x, w) + bias is emitted as ONE Conv node with weight scale*w and bias `bias` (the form every ONNX
exporter produces after BatchNorm fusion); results agree with the unfused arithmetic to fp32 rounding.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Sequence

import numpy as np

from .models import IRESNET_CFG, SCRFD_CFG

# ----------------------------------------------------------------------------------------------------- protobuf wire format
FLOAT, INT64 = 1, 7                      # TensorProto.DataType
A_FLOAT, A_INT, A_STRING, A_TENSOR, A_FLOATS, A_INTS = 1, 2, 3, 4, 6, 7   # AttributeProto.AttributeType


def _varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field: int, wire: int) -> bytes:
    return _varint((field << 3) | wire)


def _f_varint(field: int, v: int) -> bytes:
    return _key(field, 0) + _varint(int(v))


def _f_bytes(field: int, b: bytes) -> bytes:
    return _key(field, 2) + _varint(len(b)) + b


def _f_str(field: int, s: str) -> bytes:
    return _f_bytes(field, s.encode("utf-8"))


def _f_float(field: int, v: float) -> bytes:
    return _key(field, 5) + struct.pack("<f", float(v))


def tensor(name: str, arr: np.ndarray) -> bytes:
    """TensorProto: dims=1, data_type=2, name=8, raw_data=9."""
    a = np.ascontiguousarray(arr)
    if a.dtype == np.int64:
        dt = INT64
    else:
        a = a.astype(np.float32)
        dt = FLOAT
    out = b"".join(_f_varint(1, d) for d in a.shape)
    return out + _f_varint(2, dt) + _f_str(8, name) + _f_bytes(9, a.tobytes())


def attr(name: str, value) -> bytes:
    """AttributeProto: name=1, f=2, i=3, s=4, floats=7, ints=8, type=20."""
    out = _f_str(1, name)
    if isinstance(value, float):
        return out + _f_float(2, value) + _f_varint(20, A_FLOAT)
    if isinstance(value, int):
        return out + _f_varint(3, value) + _f_varint(20, A_INT)
    if isinstance(value, str):
        return out + _f_bytes(4, value.encode()) + _f_varint(20, A_STRING)
    vals = list(value)
    if vals and isinstance(vals[0], float):
        return out + b"".join(_f_float(7, v) for v in vals) + _f_varint(20, A_FLOATS)
    return out + b"".join(_f_varint(8, v) for v in vals) + _f_varint(20, A_INTS)


def node(op: str, inputs: Sequence[str], outputs: Sequence[str], name: str = "", **attrs) -> bytes:
    """NodeProto: input=1, output=2, name=3, op_type=4, attribute=5."""
    out = b"".join(_f_str(1, i) for i in inputs) + b"".join(_f_str(2, o) for o in outputs)
    out += _f_str(3, name or outputs[0]) + _f_str(4, op)
    return out + b"".join(_f_bytes(5, attr(k, v)) for k, v in attrs.items())


def value_info(name: str, shape: Sequence[int]) -> bytes:
    """ValueInfoProto{name=1, type=2{tensor_type=1{elem_type=1, shape=2{dim=1{dim_value=1}}}}}."""
    dims = b"".join(_f_bytes(1, _f_varint(1, d)) for d in shape)
    ttype = _f_varint(1, FLOAT) + _f_bytes(2, dims)
    return _f_str(1, name) + _f_bytes(2, _f_bytes(1, ttype))


def model(nodes: List[bytes], inits: List[bytes], inputs: List[bytes], outputs: List[bytes], name: str, opset: int = 13) -> bytes:
    """ModelProto{ir_version=1, producer_name=2, graph=7, opset_import=8}; GraphProto{node=1, name=2, initializer=5, input=11, output=12}."""
    g = b"".join(_f_bytes(1, n) for n in nodes) + _f_str(2, name) + b"".join(_f_bytes(5, t) for t in inits)
    g += b"".join(_f_bytes(11, i) for i in inputs) + b"".join(_f_bytes(12, o) for o in outputs)
    return _f_varint(1, 7) + _f_str(2, "person_capture_b200.oracle") + _f_bytes(7, g) + _f_bytes(8, _f_str(1, "") + _f_varint(2, opset))


# ----------------------------------------------------------------------------------------------------- graph builder
class _G:
    def __init__(self, params: Dict[str, np.ndarray]):
        self.p = {k: np.asarray(v) for k, v in params.items()}
        self.nodes: List[bytes] = []
        self.inits: List[bytes] = []
        self.n = 0

    def tmp(self, base: str) -> str:
        self.n += 1
        return f"{base}_{self.n}"

    def const(self, name: str, arr: np.ndarray) -> str:
        self.inits.append(tensor(name, arr))
        return name

    def conv(self, x: str, name: str, stride: int = 1, act: str = "none", residual: str = None) -> str:
        w = self.p[name + ".w"].astype(np.float32)
        sc = self.p[name + ".scale"].astype(np.float32).reshape(-1, 1, 1, 1)
        k = int(w.shape[-1])
        wn = self.const(name + ".W", w * sc)
        bn = self.const(name + ".B", self.p[name + ".bias"].astype(np.float32))
        y = self.tmp(name + ".y")
        self.nodes.append(node("Conv", [x, wn, bn], [y], kernel_shape=[k, k], strides=[stride, stride], pads=[k // 2] * 4,
                               dilations=[1, 1], group=1))
        if residual is not None:
            z = self.tmp(name + ".sum")
            self.nodes.append(node("Add", [y, residual], [z]))
            y = z
        if act == "relu":
            z = self.tmp(name + ".relu")
            self.nodes.append(node("Relu", [y], [z]))
            y = z
        elif act == "prelu":
            sl = self.const(name + ".slope", self.p[name + ".slope"].astype(np.float32).reshape(-1, 1, 1))
            z = self.tmp(name + ".prelu")
            self.nodes.append(node("PRelu", [y, sl], [z]))
            y = z
        return y

    def affine(self, x: str, name: str) -> str:
        s = self.const(name + ".scale", self.p[name + ".scale"].astype(np.float32).reshape(1, -1, 1, 1))
        b = self.const(name + ".bias", self.p[name + ".bias"].astype(np.float32).reshape(1, -1, 1, 1))
        m, y = self.tmp(name + ".mul"), self.tmp(name + ".aff")
        self.nodes.append(node("Mul", [x, s], [m]))
        self.nodes.append(node("Add", [m, b], [y]))
        return y

    def simple(self, op: str, inputs: Sequence[str], base: str, **attrs) -> str:
        y = self.tmp(base)
        self.nodes.append(node(op, list(inputs), [y], **attrs))
        return y


def export_scrfd(name: str, params: Dict[str, np.ndarray], S: int, path: str, layout: str = "insightface") -> List[str]:
    """Mirrors FoldedSCRFD.head_raw / run (oracle/models.py).  -> output names."""
    cfg = SCRFD_CFG[name]
    g = _G(params)
    x = g.conv("input.1", "stem1", 2, "relu")
    x = g.conv(x, "stem2", 1, "relu")
    x = g.conv(x, "stem3", 1, "relu")
    x = g.simple("MaxPool", [x], "pool", kernel_shape=[3, 3], strides=[2, 2], pads=[1, 1, 1, 1])
    outs = []
    cin = cfg["stem"][2]
    for si, (nb, planes) in enumerate(zip(cfg["blocks"], cfg["planes"])):
        for bi in range(nb):
            stride = 2 if (bi == 0 and si > 0) else 1
            p = f"s{si}.b{bi}"
            idt = x
            if stride != 1 or cin != planes:
                if stride != 1:
                    idt = g.simple("AveragePool", [x], p + ".avg", kernel_shape=[2, 2], strides=[2, 2], pads=[0, 0, 0, 0])
                idt = g.conv(idt, p + ".down", 1, "none")
            y = g.conv(x, p + ".conv1", stride, "relu")
            x = g.conv(y, p + ".conv2", 1, "relu", residual=idt)
            cin = planes
        outs.append(x)
    c = outs[1:]
    lat = [g.conv(t, f"lateral{i}") for i, t in enumerate(c)]
    scales = g.const("up.scales", np.array([1.0, 1.0, 2.0, 2.0], np.float32))
    roi = g.const("up.roi", np.zeros((0,), np.float32))
    for i in (2, 1):
        up = g.simple("Resize", [lat[i], roi, scales], f"up{i}", mode="nearest", coordinate_transformation_mode="asymmetric",
                      nearest_mode="floor")
        lat[i - 1] = g.simple("Add", [lat[i - 1], up], f"td{i}")
    inter = [g.conv(t, f"fpn{i}") for i, t in enumerate(lat)]
    for i in range(2):
        inter[i + 1] = g.simple("Add", [inter[i + 1], g.conv(inter[i], f"down{i}", 2)], f"bu{i}")
    feats = [inter[0], g.conv(inter[1], "pafpn0"), g.conv(inter[2], "pafpn1")]
    raws = []
    for li, f in enumerate(feats):
        t = f
        for i in range(cfg["stacked"]):
            t = _tower(g, t, f"tower{i}", li)
        raws.append(_tower(g, t, "out", li, act="none"))
    inputs = [value_info("input.1", [1, 3, S, S])]
    if layout == "raw":
        names = raws
        outputs = [value_info(n, [1, 30, S // s, S // s]) for n, s in zip(raws, (8, 16, 32))]
    else:
        # [1,30,h,w] -> [h,w,30] -> per-anchor rows: channels 0:2 scores (sigmoid), 2:10 boxes (x reg_scale), 10:30 landmarks
        rs = np.asarray(params["reg_scale"], np.float32).reshape(-1)
        sc, bb, kp = [], [], []
        for li, (r, s) in enumerate(zip(raws, (8, 16, 32))):
            hw = (S // s) * (S // s)
            t = g.simple("Transpose", [r], f"nhwc{li}", perm=[0, 2, 3, 1])
            parts = []
            for (a, b, per, nm) in ((0, 2, 1, "score"), (2, 10, 4, "bbox"), (10, 30, 10, "kps")):
                st = g.const(f"{nm}{li}.starts", np.array([a], np.int64))
                en = g.const(f"{nm}{li}.ends", np.array([b], np.int64))
                ax = g.const(f"{nm}{li}.axes", np.array([3], np.int64))
                sl = g.simple("Slice", [t, st, en, ax], f"{nm}{li}.slice")
                shp = g.const(f"{nm}{li}.shape", np.array([hw * 2, per], np.int64))
                parts.append(g.simple("Reshape", [sl, shp], f"{nm}{li}.rows"))
            sc.append(g.simple("Sigmoid", [parts[0]], f"score_{s}"))
            k = g.const(f"bbox{li}.k", np.array([rs[li]], np.float32))
            bb.append(g.simple("Mul", [parts[1], k], f"bbox_{s}"))
            kp.append(g.simple("Identity", [parts[2]], f"kps_{s}"))
        names = sc + bb + kp
        outputs = []
        for n, per in zip(names, [1] * 3 + [4] * 3 + [10] * 3):
            s = int(n.split("_")[1])
            outputs.append(value_info(n, [(S // s) * (S // s) * 2, per]))
    with open(path, "wb") as fh:
        fh.write(model(g.nodes, g.inits, inputs, outputs, name))
    return names


def _tower(g: _G, x: str, pname: str, level: int, act: str = "relu") -> str:
    """Head convs share their weights across the three levels: the initialisers are emitted once, the nodes per level."""
    key = pname + ".W"
    if not any(key.encode() in t for t in g.inits):
        w = g.p[pname + ".w"].astype(np.float32) * g.p[pname + ".scale"].astype(np.float32).reshape(-1, 1, 1, 1)
        g.const(pname + ".W", w)
        g.const(pname + ".B", g.p[pname + ".bias"].astype(np.float32))
    k = int(g.p[pname + ".w"].shape[-1])
    y = g.tmp(f"{pname}.l{level}")
    g.nodes.append(node("Conv", [x, pname + ".W", pname + ".B"], [y], kernel_shape=[k, k], strides=[1, 1], pads=[k // 2] * 4,
                        dilations=[1, 1], group=1))
    if act == "relu":
        z = g.tmp(f"{pname}.l{level}.relu")
        g.nodes.append(node("Relu", [y], [z]))
        y = z
    return y


def export_iresnet(name: str, params: Dict[str, np.ndarray], path: str, batch: int = 1) -> str:
    """Mirrors FoldedIResNet._fwd (oracle/models.py).  -> output name."""
    g = _G(params)
    x = g.conv("input.1", "stem", 1, "prelu")
    for si, nb in enumerate(IRESNET_CFG[name]):
        for bi in range(nb):
            p = f"s{si}.b{bi}"
            stride = 2 if bi == 0 else 1
            idt = g.conv(x, p + ".down", stride, "none") if bi == 0 else x
            y = g.affine(x, p + ".bn1")
            y = g.conv(y, p + ".conv1", 1, "prelu")
            x = g.conv(y, p + ".conv2", stride, "none", residual=idt)
    x = g.affine(x, "bn2")
    # the FC weight is stored for an (h, w, c) flatten (oracle/models.py): NCHW -> NHWC -> [n, 7*7*512] -> Gemm
    t = g.simple("Transpose", [x], "nhwc", perm=[0, 2, 3, 1])
    shp = g.const("flat.shape", np.array([batch, -1], np.int64))
    f = g.simple("Reshape", [t, shp], "flat")
    w = g.p["fc.w"].astype(np.float32) * g.p["fc.scale"].astype(np.float32).reshape(-1, 1)
    wn = g.const("fc.W", w)
    bn = g.const("fc.B", g.p["fc.bias"].astype(np.float32).reshape(-1))
    g.nodes.append(node("Gemm", [f, wn, bn], ["embedding"], alpha=1.0, beta=1.0, transA=0, transB=1))
    with open(path, "wb") as fh:
        fh.write(model(g.nodes, g.inits, [value_info("input.1", [batch, 3, 112, 112])], [value_info("embedding", [batch, 512])], name))
    return "embedding"


if __name__ == "__main__":
    # python -m oracle.onnx_export scrfd_10g_bnkps 640 out.onnx | python -m oracle.onnx_export arcface_r100 out.onnx
    import sys

    from person_capture_b200 import weights

    model_name = sys.argv[1]
    if model_name.startswith("scrfd"):
        print(export_scrfd(model_name, weights.load_params(model_name), int(sys.argv[2]), sys.argv[3]))
    else:
        print(export_iresnet(model_name, weights.load_params(model_name), sys.argv[2]))
