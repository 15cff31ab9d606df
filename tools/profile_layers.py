"""Per-layer timing of conv_tc_kernel (CUDA events, PCB_PROFILE_DUMP): SCRFD at S x S for a batch of frames
and ArcFace for a batch of chips.  usage: python tools/profile_layers.py [--scrfd scrfd_10g_bnkps --S 512 --n 64 --faces 128]"""
import argparse, collections, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--scrfd", default="scrfd_10g_bnkps")
ap.add_argument("--arcface", default="arcface_r100")
ap.add_argument("--S", type=int, default=512)
ap.add_argument("--n", type=int, default=64)
ap.add_argument("--faces", type=int, default=128)
ap.add_argument("--out", default="gpurun_out/layers.csv")
ap.add_argument("--impl", type=int, default=0)
ap.add_argument("--reps", type=int, default=1, help="profiled passes (times are summed; divide by reps)")
ap.add_argument("--sustain", type=int, default=0, help="run the ArcFace pass this many times back to back and report TF/s + clocks")
args = ap.parse_args()
if os.path.exists(args.out):
    os.remove(args.out)
os.environ["PCB_PROFILE_DUMP"] = args.out
import torch
from person_capture_b200.engine import Engine
eng = Engine(0, scrfd=args.scrfd, arcface=args.arcface)
eng.set_conv_impl(args.impl)
rng = np.random.default_rng(0)
frames = eng.to_device(rng.integers(0, 256, (args.n, 540, 960, 3), dtype=np.uint8))
chips = eng.to_device(rng.integers(0, 256, (args.faces, 112, 112, 3), dtype=np.uint8))
for it in range(3):
    eng.detect(frames, args.S, 0.5)
    eng.embed(chips, args.faces, True)
eng.sync()
eng.profile_read(reset=True)
if os.path.exists(args.out):
    os.remove(args.out)
eng.set_profile(True)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
for rep in range(args.reps):
    t0.record(eng.stream)
    eng.detect(frames, args.S, 0.5)
    t1.record(eng.stream)
    eng.embed(chips, args.faces, True)
    t2.record(eng.stream)
ms, fl, n = eng.profile_read()
print(f"scrfd pass {t0.elapsed_time(t1):.3f} ms for {args.n} frames; arcface pass {t1.elapsed_time(t2):.3f} ms for {2*args.faces} images (last of {args.reps} passes)")
print(f"conv total {ms:.3f} ms, {fl/1e12:.3f} TFLOP -> {fl/ms/1e9:.1f} TFLOP/s over {n} launches")
rows = collections.OrderedDict()
for line in open(args.out):
    parts = line.strip().split(",")
    key = ",".join(parts[:-2]); t = float(parts[-2]); f = float(parts[-1])
    a = rows.setdefault(key, [0, 0.0, 0.0]); a[0] += 1; a[1] += t; a[2] += f
print(f"{'layer shape':110s} {'cnt':>4s} {'ms':>9s} {'TF/s':>8s} {'%time':>6s}")
for k, (c, t, f) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:110s} {c:4d} {t:9.3f} {f/t/1e9:8.1f} {100*t/ms:6.1f}")

if args.sustain:
    import subprocess, threading, time
    from person_capture_b200 import graphs
    eng.set_profile(False)
    rows = []
    pr = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "50"],
                          stdout=subprocess.PIPE, text=True)
    threading.Thread(target=lambda: [rows.append(l.strip()) for l in pr.stdout], daemon=True).start()
    time.sleep(0.3)
    n0 = len(rows)
    a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
    a0.record(eng.stream)
    for _ in range(args.sustain):
        eng.embed(chips, args.faces, True)
    a1.record(eng.stream)
    eng.sync()
    n1 = len(rows)
    pr.terminate()
    ms = a0.elapsed_time(a1)
    fl = 2.0 * graphs.graph_macs(graphs.build_graph(args.arcface), 112, 112) * 2 * args.faces * args.sustain
    clk = [float(r.split(",")[0]) for r in rows[n0:n1] if r]
    pw = [float(r.split(",")[1]) for r in rows[n0:n1] if r]
    print(f"sustained ArcFace: {args.sustain} passes of {2*args.faces} images in {ms:.1f} ms -> {fl/ms/1e9:.1f} TFLOP/s (whole pass incl. non-conv kernels); "
          f"sm clock median {np.median(clk) if clk else -1:.0f} MHz min {min(clk) if clk else -1:.0f}, power median {np.median(pw) if pw else -1:.0f} W, samples {len(clk)}")
