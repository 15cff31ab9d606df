"""Weight files for the four graphs on the path ("folded" format).

The reference downloads trained ONNX files at run time (person_capture/face_embedder.py:55-83);
there is no network here, so:
  * SCRFD-10G / SCRFD-2.5G: small detectors trained on the synthetic faces of `synth.py`
    (tooling: oracle/train_scrfd.py), committed as fp16 .npz under weights/.
  * ArcFace iResNet-50 / -100: 44 M / 65 M parameters cannot be committed, so conv/FC
    weights are regenerated from a fixed numpy PCG64 seed (bit-identical on every machine)
    and only the calibrated per-channel affine terms (weights/arcface_*_affine.npz, written
    by oracle/calibrate_arcface.py) are committed.
Both the CUDA path and the CPU oracle load parameters through `load_params`, so they consume
the same numbers.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Tuple

import numpy as np

WEIGHTS_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights")

IRESNET_BLOCKS = {"arcface_r50": (3, 4, 14, 3), "arcface_r100": (3, 13, 30, 3)}
ARCFACE_SEED = 20240


def iresnet_convs(name: str) -> List[Tuple[str, int, int, int]]:
    """[(conv name, cin, cout, k)] in the order weights are drawn from the RNG."""
    out = [("stem", 3, 64, 3)]
    cin = 64
    for si, nb in enumerate(IRESNET_BLOCKS[name]):
        planes = 64 << si
        for bi in range(nb):
            p = f"s{si}.b{bi}"
            if bi == 0:
                out.append((p + ".down", cin, planes, 1))
            out.append((p + ".conv1", cin, planes, 3))
            out.append((p + ".conv2", planes, planes, 3))
            cin = planes
    return out


def arcface_random_weights(name: str, seed: int = ARCFACE_SEED) -> Dict[str, np.ndarray]:
    """He-init conv weights + FC weight [512, 7*7*512] ((h,w,c) flatten order), fp16."""
    rng = np.random.default_rng(seed + (0 if name.endswith("r100") else 1))
    P: Dict[str, np.ndarray] = {}
    for cname, cin, cout, k in iresnet_convs(name):
        std = math.sqrt(2.0 / (cin * k * k))
        P[cname + ".w"] = (rng.standard_normal((cout, cin, k, k), dtype=np.float32) * std).astype(np.float16)
    kfc = 7 * 7 * 512
    P["fc.w"] = (rng.standard_normal((512, kfc), dtype=np.float32) * math.sqrt(1.0 / kfc)).astype(np.float16)
    return P


_CACHE: Dict[str, Dict[str, np.ndarray]] = {}


def load_params(name: str) -> Dict[str, np.ndarray]:
    """Folded parameter dict for `name` in {scrfd_10g_bnkps, scrfd_2.5g_bnkps, arcface_r50, arcface_r100}."""
    if name in _CACHE:
        return _CACHE[name]
    if name.startswith("scrfd"):
        path = os.path.join(WEIGHTS_DIR, name + ".npz")
        if not os.path.isfile(path):
            raise FileNotFoundError(f"{path} missing (train with oracle/train_scrfd.py)")
        with np.load(path) as z:
            P = {k: z[k] for k in z.files}
    elif name.startswith("arcface"):
        path = os.path.join(WEIGHTS_DIR, name + "_affine.npz")
        if not os.path.isfile(path):
            raise FileNotFoundError(f"{path} missing (run oracle/calibrate_arcface.py)")
        P = arcface_random_weights(name)
        with np.load(path) as z:
            for k in z.files:
                P[k] = z[k]
    else:
        raise ValueError(name)
    _CACHE[name] = P
    return P
