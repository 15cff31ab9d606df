"""ORACLE (test infrastructure): the identity calls of the reference's dataset curator, restated on the CPU.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this package.

person_capture/dataset_curator.py: reference feature from the top-quality face of the ref image (:343-354), centred
square letterbox to 640 (:384-401), best face + box un-mapping (:404-426), 1-row `_fd_min` (:617-627) and the identity
fields of `describe` (:629-648).  Scoring / MMR selection / pHash are downstream of identity and out of scope."""
from __future__ import annotations

from typing import Optional

import cv2
import numpy as np


def letterbox_square(bgr: np.ndarray, size: int = 640):
    height, width = bgr.shape[:2]
    if height == 0 or width == 0:
        return np.zeros((size, size, 3), np.uint8), 1.0, 0, 0
    scale = min(size / float(width), size / float(height))
    new_w, new_h = int(round(width * scale)), int(round(height * scale))
    resized = cv2.resize(bgr, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
    canvas = np.zeros((size, size, 3), bgr.dtype)
    dx, dy = (size - new_w) // 2, (size - new_h) // 2
    canvas[dy:dy + new_h, dx:dx + new_w] = resized
    return canvas, float(scale), int(dx), int(dy)


def best_face(faces):
    return max(faces, key=lambda f: (f["quality"], (f["bbox"][2] - f["bbox"][0]) * (f["bbox"][3] - f["bbox"][1]))) if faces else None


class CuratorIdentityOracle:
    def __init__(self, face, ref_bgr: Optional[np.ndarray], det_square: bool = True, id_already_passed: bool = False):
        self.face, self.det_square, self.id_already_passed = face, det_square, id_already_passed
        self.ref_feat = None
        if ref_bgr is not None:
            rfaces = face.extract(ref_bgr)
            if rfaces:
                self.ref_feat = max(rfaces, key=lambda f: f.get("quality", 0.0)).get("feat")

    def fd_min(self, feat) -> float:
        if self.id_already_passed:
            return 0.0
        if feat is None or self.ref_feat is None:
            return 9.0
        v = np.asarray(feat, np.float32)
        v = v / max(1e-6, float(np.linalg.norm(v)))
        r = np.asarray(self.ref_feat, np.float32)
        r = r / max(1e-6, float(np.linalg.norm(r)))
        return float(1.0 - float(np.dot(v, r)))

    def detect_best_face(self, bgr):
        height, width = bgr.shape[:2]
        if self.det_square:
            canvas, scale, dx, dy = letterbox_square(bgr, 640)
            best = best_face(self.face.extract(canvas))
            if best is None:
                return None
            x1, y1, x2, y2 = [float(v) for v in best["bbox"]]
            inv = 1.0 / max(scale, 1e-6)
            ox1 = max(0, min(width, int(round((x1 - dx) * inv))))
            oy1 = max(0, min(height, int(round((y1 - dy) * inv))))
            ox2 = max(ox1 + 1, min(width, int(round((x2 - dx) * inv))))
            oy2 = max(oy1 + 1, min(height, int(round((y2 - dy) * inv))))
            out = dict(best)
            out["bbox"] = (int(ox1), int(oy1), int(ox2), int(oy2))
            return out
        return best_face(self.face.extract(bgr))

    def describe(self, bgr):
        """-> dict(bbox, fd, quality, feat) -- the identity fields of Curator.describe."""
        best = self.detect_best_face(bgr)
        fd = 0.0 if self.id_already_passed else 9.0
        if best is None:
            return dict(bbox=None, fd=fd, quality=0.0, feat=None)
        if not self.id_already_passed:
            fd = self.fd_min(best.get("feat"))
        return dict(bbox=tuple(int(x) for x in best["bbox"]), fd=fd, quality=float(best.get("quality", 0.0)), feat=best.get("feat"))
