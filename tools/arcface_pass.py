"""One ArcFace pass (chip patch -> iResNet graph -> K5 match) over a production-size run, bracketed by cudaProfilerStart/Stop so
that `ncu --profile-from-start off ...` sees exactly its launches.  Conv launch order (iResNet-100): 0 stem | stage 1: 1-7 |
stage 2: 8-34 | stage 3: 35 down, 36 conv1 28x28, 37 conv2 stride 2, then per block conv1 / conv2 at 14x14 (38, 39, ...).
usage: python tools/arcface_pass.py [--arcface arcface_r100 --images 504]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--arcface", default="arcface_r100")
ap.add_argument("--images", type=int, default=504)
ap.add_argument("--reps", type=int, default=4)
args = ap.parse_args()
import torch
from person_capture_b200.engine import Engine
eng = Engine(0, scrfd=None, arcface=args.arcface)
rng = np.random.default_rng(0)
chips = eng.to_device(rng.integers(0, 256, (args.images, 112, 112, 3), dtype=np.uint8))
bank = rng.normal(size=(64, 512)).astype(np.float32)
eng.set_bank(bank / np.linalg.norm(bank, axis=1, keepdims=True))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for r in range(args.reps):
    if r == args.reps - 1:
        torch.cuda.profiler.start()
    e0.record(eng.stream)
    emb, _ = eng.embed(chips, args.images, False)
    eng.match(emb, None, None, args.images)
    e1.record(eng.stream)
    eng.sync()
    ts.append(e0.elapsed_time(e1))
torch.cuda.profiler.stop()
print(f"{args.arcface} {args.images} images: {np.median(ts):.3f} ms per pass (median of {args.reps})")
