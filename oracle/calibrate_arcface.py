"""ORACLE TOOLING: compute the per-layer affine terms of the seeded-random ArcFace graphs.

usage: python -m oracle.calibrate_arcface arcface_r50 arcface_r100
Writes weights/<name>_affine.npz (small: scale/bias/slope vectors only).  The conv/FC weights
themselves are regenerated from the seed by person_capture_b200/weights.py on every machine.
Calibration chips are synthetic aligned faces (person_capture_b200/synth.py) with jitter.
"""
from __future__ import annotations

import os
import sys
import time

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.models import calibrate_arcface_affine  # noqa: E402
from person_capture_b200 import synth, weights  # noqa: E402


def calib_chips(n: int, seed: int = 4242) -> np.ndarray:
    rng = np.random.default_rng(seed)
    X = np.empty((n, 3, 112, 112), np.float32)
    for i in range(n):
        canvas = synth.background(rng, 160, 160, clutter=3)
        side = float(rng.uniform(100, 125))
        synth.paste_face(canvas, 5000 + i, 80 + rng.uniform(-3, 3), 80 + rng.uniform(-3, 3), side,
                         float(rng.uniform(-6, 6)), float(rng.uniform(0.85, 1.15)))
        chip = canvas[24:136, 24:136]
        rgb = cv2.cvtColor(chip, cv2.COLOR_BGR2RGB)
        X[i] = np.transpose(rgb.astype(np.float32) / 127.5 - 1.0, (2, 0, 1))
    return X


def main():
    torch.set_num_threads(int(os.environ.get("CALIB_THREADS", "4")))
    for name in sys.argv[1:]:
        t0 = time.time()
        W = weights.arcface_random_weights(name)
        P = calibrate_arcface_affine(name, W, calib_chips(40))
        out = os.path.join(weights.WEIGHTS_DIR, name + "_affine.npz")
        np.savez(out, **P)
        print(name, "saved", out, f"{time.time() - t0:.0f}s", sum(v.nbytes for v in P.values()), "bytes")


if __name__ == "__main__":
    main()
