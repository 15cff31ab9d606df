"""One SCRFD pass (K1 letterbox -> graph -> K3 decode/NMS -> K4 align) over a frame batch, bracketed by cudaProfilerStart/Stop so
that `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv` lists exactly its launches.
usage: python tools/scrfd_pass.py [--S 512 --n 64 --scrfd scrfd_10g_bnkps] ; prints the CUDA-event time of the pass."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--scrfd", default="scrfd_10g_bnkps")
ap.add_argument("--S", type=int, default=512)
ap.add_argument("--n", type=int, default=64)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
import torch
from person_capture_b200 import synth
from person_capture_b200.engine import Engine
eng = Engine(0, scrfd=args.scrfd, arcface=None)
clip = synth.ClipSpec(960, 540, 8, seed=1002, distractor_prob=1.0, target_segments=[(0, 7)])
frames = eng.to_device(np.stack([clip.frame(i % 8) for i in range(args.n)]))
# K0 input: full-resolution frames for the INTER_AREA downscale of the pre-scan (1080p -> 960x540)
full = eng.to_device(np.random.default_rng(0).integers(0, 256, (args.n, 1080, 1920, 3), dtype=np.uint8))
for _ in range(3):
    det = eng.detect(frames, args.S, 0.5)
    eng.align(frames, det, max_faces=4096)
eng.sync()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
ts = []
for r in range(args.reps):
    if r == args.reps - 1:
        torch.cuda.profiler.start()
    small = eng.resize(full, 540, 960, area=True)          # K0 (outside the detect timing below)
    e0.record(eng.stream)
    det = eng.detect(frames, args.S, 0.5)
    e1.record(eng.stream)
    eng.align(frames, det, max_faces=4096)
    e2.record(eng.stream)
    eng.sync()
    ts.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
torch.cuda.profiler.stop()
print(f"{args.scrfd} S={args.S} n={args.n}: detect {np.median([t[0] for t in ts]):.3f} ms, align {np.median([t[1] for t in ts]):.3f} ms (median of {args.reps}); "
      f"launches per pass: {eng.launch_count() // (args.reps + 3)}")
