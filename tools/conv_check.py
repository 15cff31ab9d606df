"""Layer-by-layer agreement of the tensor-core conv kernels with the CUDA-core validation kernel.
Runs each graph once per implementation on the same input and reports, per op, the worst absolute and
relative difference of its output tensor.  usage: python tools/conv_check.py [--impl 0] [--n 2] [--S 320]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--impl", type=int, default=0)
ap.add_argument("--n", type=int, default=2)
ap.add_argument("--faces", type=int, default=3)
ap.add_argument("--S", type=int, default=320)
ap.add_argument("--scrfd", default="scrfd_10g_bnkps")
ap.add_argument("--arcface", default="arcface_r50")
ap.add_argument("--verbose", action="store_true")
args = ap.parse_args()
from person_capture_b200.engine import Engine
from person_capture_b200 import _lib as L
from person_capture_b200 import _lib as _L
eng = Engine(0, scrfd=args.scrfd, arcface=args.arcface, lib_path=_L.VAL_LIB_PATH)   # the validation kernel lives in the test-only library
rng = np.random.default_rng(0)
import cv2
frames = np.stack([cv2.GaussianBlur(rng.integers(0, 256, (270, 480, 3), dtype=np.uint8), (0, 0), 1.5) for _ in range(args.n)])
chips = np.stack([cv2.GaussianBlur(rng.integers(0, 256, (112, 112, 3), dtype=np.uint8), (0, 0), 1.5) for _ in range(args.faces)])
fd, cd = eng.to_device(frames), eng.to_device(chips)


def run(slot, impl):
    eng.set_conv_impl(impl)
    if slot == L.MODEL_SCRFD:
        eng.detect(fd, args.S, 0.5)
    else:
        eng.embed(cd, args.faces, True)
    eng.sync()
    g = eng.graphs[slot]
    out = {}
    for op in g.ops:
        for t in (op["out"], op.get("out2", -1)):
            if t is not None and t >= 0:
                out[t] = eng.get_tensor(slot, t)
    eng.set_conv_impl(0)
    return out


bad = 0
for slot, name in ((L.MODEL_SCRFD, args.scrfd), (L.MODEL_ARCFACE, args.arcface)):
    ref = run(slot, 1)
    got = run(slot, args.impl)
    g = eng.graphs[slot]
    worst = 0.0
    for op in g.ops:
        for t in (op["out"], op.get("out2", -1)):
            if t is None or t < 0:
                continue
            a, b = ref[t], got[t]
            d = float(np.abs(a - b).max())
            s = float(np.abs(a).max()) + 1e-6
            rel = d / s
            worst = max(worst, rel)
            flag = "BAD" if (rel > 2e-2 or not np.isfinite(b).all()) else ""
            if flag:
                bad += 1
            if args.verbose or flag:
                print(f"{name} op kind={op['kind']} k={op['k']} cin={op['cin']} cout={op['cout']} stride={op['stride']} tensor={t} "
                      f"shape={a.shape} maxabs={d:.4g} scale={s:.4g} rel={rel:.3g} {flag}")
    print(f"{name}: impl {args.impl} vs validation kernel: worst relative diff {worst:.3g}")
print("CONV_CHECK", "FAIL" if bad else "OK", f"pair={os.environ.get('PCB_CONV_PAIR','1')} iss={os.environ.get('PCB_CONV_ISS','2')} mt={os.environ.get('PCB_CONV_MT','auto')}")
sys.exit(1 if bad else 0)
