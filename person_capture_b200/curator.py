"""Identity side of the reference's dataset curator on the GPU path (a second consumer of the FaceEmbedder boundary).

Reproduces person_capture/dataset_curator.py: reference feature = top-quality face of the ref image (:343-354); per crop a
centred square letterbox to 640 (:384-401, INTER_LINEAR -- done by pcb_resize_linear, bit-exact with cv2), FaceEmbedder.extract,
best face, box un-mapping (:404-426); 1-row `_fd_min` (:617-627; computed by pcb_match with the reference as a one-row bank).
Scoring, MMR selection, pHash and CLIP diversity are downstream product logic and out of scope."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .face_embedder import FaceEmbedder


class CuratorIdentity:
    def __init__(self, face: FaceEmbedder, ref_bgr: Optional[np.ndarray], det_square: bool = True, id_already_passed: bool = False):
        self.face, self.det_square, self.id_already_passed = face, det_square, id_already_passed
        self.ref_feat = None
        if ref_bgr is not None:
            rfaces = face.extract(ref_bgr)
            if rfaces:
                self.ref_feat = max(rfaces, key=lambda f: f.get("quality", 0.0)).get("feat")

    def letterbox_square(self, bgr: np.ndarray, size: int = 640):
        """-> (device canvas uint8 [size,size,3], scale, dx, dy); the resize runs on the GPU."""
        eng = self.face.engine
        height, width = bgr.shape[:2]
        scale = min(size / float(width), size / float(height))
        new_w, new_h = int(round(width * scale)), int(round(height * scale))
        src = eng.to_device(bgr[None])
        resized = eng.resize(src, new_h, new_w, area=False)
        dx, dy = (size - new_w) // 2, (size - new_h) // 2
        with torch.cuda.stream(eng.stream):
            canvas = torch.zeros((size, size, 3), dtype=torch.uint8, device=eng.tdev)
            canvas[dy:dy + new_h, dx:dx + new_w] = resized[0]
        return canvas, float(scale), int(dx), int(dy)

    def fd_min(self, feat_dev: torch.Tensor, count: int) -> np.ndarray:
        """fd of the `count` device-resident features against the one-row reference (pcb_match renormalises both sides)."""
        eng = self.face.engine
        r = np.asarray(self.ref_feat, np.float32)
        r = (r / max(1e-6, float(np.linalg.norm(r)))).reshape(1, -1)
        eng.set_bank(r)
        _, sim, _ = eng.match(feat_dev, None, None, count, want_feat=False)
        eng.sync()
        return 1.0 - sim[:count].cpu().numpy().astype(np.float64)

    def describe(self, bgr: np.ndarray):
        """-> dict(bbox, fd, quality, feat): the identity fields of Curator.describe for one saved crop."""
        height, width = bgr.shape[:2]
        fd = 0.0 if self.id_already_passed else 9.0
        if height == 0 or width == 0:
            return dict(bbox=None, fd=fd, quality=0.0, feat=None)
        if self.det_square:
            canvas, scale, dx, dy = self.letterbox_square(bgr, 640)
            faces = self.face.extract(canvas)
        else:
            scale, dx, dy = 1.0, 0, 0
            faces = self.face.extract(bgr)
        if not faces:
            return dict(bbox=None, fd=fd, quality=0.0, feat=None)
        best = faces[0]                      # extract() returns faces sorted by (quality, area): element 0 is best_face
        x1, y1, x2, y2 = [float(v) for v in best["bbox"]]
        if self.det_square:
            inv = 1.0 / max(scale, 1e-6)
            ox1 = max(0, min(width, int(round((x1 - dx) * inv))))
            oy1 = max(0, min(height, int(round((y1 - dy) * inv))))
            bbox = (ox1, oy1, max(ox1 + 1, min(width, int(round((x2 - dx) * inv)))), max(oy1 + 1, min(height, int(round((y2 - dy) * inv)))))
        else:
            bbox = tuple(int(v) for v in best["bbox"])
        if not self.id_already_passed and self.ref_feat is not None:
            fd = float(self.fd_min(self.face.last_feats_dev, 1)[0])
        return dict(bbox=bbox, fd=fd, quality=float(best["quality"]), feat=best["feat"])
