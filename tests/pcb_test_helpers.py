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
