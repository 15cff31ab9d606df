"""HBM-bound stages of the path against the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs): algorithmic bytes of each
call / CUDA-event time on the engine's stream, inputs larger than L2 or rotated between iterations.
usage: python tools/hbm_kernels.py [--iters 20]"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
import torch
from person_capture_b200.engine import Engine
from person_capture_b200 import _lib as L

peak = 6547.2
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
eng = Engine(0, scrfd="scrfd_10g_bnkps", arcface="arcface_r100")
rng = np.random.default_rng(0)


def timed(fn, nbytes, name, sets):
    """sets: list of argument tuples rotated between iterations (each set larger than or disjoint in L2)."""
    for s in sets:
        fn(*s)
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.stream)
    for i in range(args.iters):
        fn(*sets[i % len(sets)])
    e1.record(eng.stream)
    eng.sync()
    ms = e0.elapsed_time(e1) / args.iters
    gbs = nbytes / ms / 1e6
    print(f"{name:58s} {ms*1000:9.1f} us  {nbytes/1e6:9.1f} MB  {gbs:8.1f} GB/s  {100*gbs/peak:5.1f}% of measured HBM copy peak")


# K0: exact 2x INTER_AREA, 64 x 1080p -> 960x540 (398 MB in, 99.5 MB out: > L2)
n = 64
frames = [eng.to_device(rng.integers(0, 256, (n, 1080, 1920, 3), dtype=np.uint8)) for _ in range(2)]
outs = [eng.empty((n, 540, 960, 3), torch.uint8) for _ in range(2)]
timed(lambda a, o: eng._check(eng.lib.pcb_resize_area(eng.ctx, a.data_ptr(), n, 1080, 1920, o.data_ptr(), 540, 960), "k0"),
      n * (1080 * 1920 * 3 + 540 * 960 * 3), "K0 area2x 64x1080p -> 960x540", list(zip(frames, outs)))
# K1: letterbox + normalise + patch tensor, 64 x 960x540 -> S=512
small = [eng.resize(f, 540, 960, area=True) for f in frames]
pt = [eng.zeros((n, 258, 258, 32), torch.float16) for _ in range(2)]
timed(lambda a, o: eng._check(eng.lib.pcb_letterbox(eng.ctx, a.data_ptr(), n, 540, 960, 512, 0, 0, o.data_ptr(), None), "k1"),
      n * (540 * 960 * 3 + 256 * 144 * 32 * 2), "K1 letterbox 64x(960x540) -> 512^2 patch tensor (content rows)", list(zip(small, pt)))
del frames, outs
# ArcFace input: chips -> patch tensor (444 faces, plain)
f = 444
chips = [eng.to_device(rng.integers(0, 256, (f, 112, 112, 3), dtype=np.uint8)) for _ in range(3)]
emb = eng.empty((f, 512), torch.float32)
# (chip_patch is internal to pcb_embed; time the stand-alone K5 and pooling ops instead)
# K5: match 4096 features against a 10 000-row bank (config 4)
bank = rng.normal(size=(10000, 512)).astype(np.float32)
bank /= np.linalg.norm(bank, axis=1, keepdims=True)
eng.set_bank(bank)
F = 4096
feats = [eng.to_device(rng.normal(size=(F, 512)).astype(np.float32)) for _ in range(2)]
sim = eng.empty((F,), torch.float32); arg = eng.empty((F,), torch.int32)
timed(lambda a: eng._check(eng.lib.pcb_match(eng.ctx, a.data_ptr(), None, None, F, None, sim.data_ptr(), arg.data_ptr()), "k5"),
      10000 * 512 * 4 + F * 512 * 4, "K5 match 4096 faces x 10k bank (bank read once = algorithmic)", [(x,) for x in feats])
print("note: K5 re-reads the 20 MB bank from L2 for every face block; its DRAM traffic is the algorithmic figure, its time is L2/FMA bound")
# K3+K4 are latency-bound (a few hundred faces): report times only
det = eng.detect(small[0], 512, 0.5)
eng.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(eng.stream)
for _ in range(args.iters):
    al = eng.align(small[0], det, max_faces=4096)
e1.record(eng.stream)
eng.sync()
print(f"K4 cross-pass NMS + LMedS + warp + quality, 64 frames / {int(al.face_total.cpu()[0])} faces: {e0.elapsed_time(e1)/args.iters*1000:.1f} us per call")
