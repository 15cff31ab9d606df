"""Shared helpers for the parity tests: oracle construction and comparison utilities."""
import numpy as np

from oracle.models import FoldedIResNet, FoldedSCRFD
from oracle.scrfd_detect import SCRFDOracle
from oracle.face_embedder import FaceEmbedderOracle
from person_capture_b200 import weights

_NETS = {}


def oracle_scrfd(name):
    if name not in _NETS:
        _NETS[name] = FoldedSCRFD(name, weights.load_params(name))
    return _NETS[name]


def oracle_arcface(name):
    if name not in _NETS:
        _NETS[name] = FoldedIResNet(name, weights.load_params(name))
    return _NETS[name]


def oracle_embedder(scrfd="scrfd_10g_bnkps", arcface="arcface_r50", conf=0.30):
    return FaceEmbedderOracle(SCRFDOracle(oracle_scrfd(scrfd)), oracle_arcface(arcface), conf=conf)


def cos(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))


def smooth_image(rng, h, w):
    import cv2
    a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    return np.ascontiguousarray(cv2.GaussianBlur(a, (0, 0), 1.0))


_PROJ = None


def proj_arcface(x):
    """Stand-in ArcFace session for tests that pin host logic against the reference (tests/golden/ref_harness.py): a fixed
    linear map of the 8x8-pooled blob, float64 accumulate.  float32 [n,3,112,112] -> float32 [n,512]."""
    global _PROJ
    if _PROJ is None:
        _PROJ = np.random.default_rng(4242).standard_normal((588, 512))
    x = np.asarray(x, np.float32)
    n = x.shape[0]
    pooled = x.reshape(n, 3, 14, 8, 14, 8).astype(np.float64).mean(axis=(3, 5)).reshape(n, 588)
    return (pooled @ _PROJ).astype(np.float32)


class ProjArcface:
    """The same map behind the oracle's `arcface.run(blob)` interface."""

    def run(self, x, batch: int = 16):
        return proj_arcface(x)
