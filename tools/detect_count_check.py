"""Detections per frame on the bench pool for each conv implementation (0 product, 2 baseline tcgen05, 1 CUDA-core)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from person_capture_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
frames, _ = bench.make_pool(n)
from person_capture_b200 import _lib as _L
eng = Engine(0, scrfd="scrfd_10g_bnkps", arcface=None, lib_path=_L.VAL_LIB_PATH)
fd = eng.resize(eng.to_device(frames), 540, 960, area=True)
res = {}
for impl in (1, 2, 0):
    eng.set_conv_impl(impl)
    for bs in (n, 8):
        counts = []
        for b0 in range(0, n, bs):
            d = eng.detect(fd[b0:b0 + bs].contiguous(), 512, 0.5)
            eng.sync()
            counts += d.acc_count.cpu().numpy().tolist()
        res[(impl, bs)] = counts
        print(f"impl {impl} batch {bs}: total {sum(counts)} per-frame {counts[:16]}")
ref = res[(1, 8)]
for k, v in res.items():
    print(k, "== validation(batch 8):", v == ref)
