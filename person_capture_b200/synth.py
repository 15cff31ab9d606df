"""Seeded synthetic clips for the identity hot path (SURVEY.md §8d "Synthetic inputs").

There is no network and no dataset in the build image, so frames, reference images and
identities are generated.  A *face* is a square patch with an identity-specific texture
and five landmark blobs placed at the ArcFace 112x112 template positions
(reference template: person_capture/face_embedder.py:1279), so that a detector trained on
these patches emits canonical landmarks and the production branch of the path
(`_canon_5pts` -> `_align_by_5pts`, face_embedder.py:1430-1473) is the one exercised.

Pure numpy + cv2; used by tests/, bench.py and the oracle tooling.  Not on the hot path.
"""
from __future__ import annotations

import numpy as np
import cv2

# ArcFace landmark template in a 112x112 chip (face_embedder.py:1279), as fractions of the box.
ARC_DST = np.array(
    [[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366], [41.5493, 92.3655], [70.7299, 92.2041]],
    dtype=np.float32,
)
_LM_FRAC = ARC_DST / 112.0

_LM_COLORS_BGR = np.array(
    [[30, 30, 30], [30, 30, 30], [235, 235, 235], [40, 40, 200], [40, 40, 200]], dtype=np.float32
)


def identity_texture(identity: int, size: int = 128) -> np.ndarray:
    """Canonical `size`x`size` BGR uint8 face patch for `identity` (deterministic)."""
    rng = np.random.default_rng(7_000_003 + 7919 * int(identity))
    low = rng.uniform(60, 220, size=(6, 6, 3)).astype(np.float32)
    base = cv2.resize(low, (size, size), interpolation=cv2.INTER_CUBIC)
    mid = rng.uniform(-38, 38, size=(16, 16, 3)).astype(np.float32)
    base += cv2.resize(mid, (size, size), interpolation=cv2.INTER_CUBIC)
    fine = rng.uniform(-22, 22, size=(48, 48, 3)).astype(np.float32)
    base += cv2.resize(fine, (size, size), interpolation=cv2.INTER_LINEAR)
    # landmark blobs: eyes dark, nose bright, mouth corners red; a thin bright ring round each
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    for k in range(5):
        cx, cy = _LM_FRAC[k] * size
        r = 0.055 * size if k != 2 else 0.048 * size
        d = np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2)
        ring = np.clip(1.0 - np.abs(d - 1.45 * r) / (0.35 * r), 0, 1)[..., None]
        base = base * (1 - 0.8 * ring) + 250.0 * 0.8 * ring
        core = np.clip((r - d) / (0.25 * r) + 1.0, 0, 1)[..., None]
        base = base * (1 - core) + _LM_COLORS_BGR[k] * core
    # dark frame so that the box outline is a learnable cue
    b = max(2, size // 32)
    base[:b, :] = 25
    base[-b:, :] = 25
    base[:, :b] = 25
    base[:, -b:] = 25
    return np.clip(base, 0, 255).astype(np.uint8)


def background(rng: np.random.Generator, h: int, w: int, clutter: int = 6) -> np.ndarray:
    """Smooth noise background plus a few rectangles/discs as distractors."""
    sh, sw = max(2, h // 24), max(2, w // 24)
    low = rng.uniform(20, 235, size=(sh, sw, 3)).astype(np.float32)
    img = cv2.resize(low, (w, h), interpolation=cv2.INTER_CUBIC)
    img += cv2.GaussianBlur(rng.uniform(-30, 30, size=(h, w, 3)).astype(np.float32), (0, 0), 1.2)
    for _ in range(clutter):
        col = rng.uniform(0, 255, size=3).tolist()
        x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
        s = int(rng.integers(max(4, min(h, w) // 40), max(8, min(h, w) // 5)))
        if rng.random() < 0.5:
            cv2.rectangle(img, (x, y), (x + s, y + int(s * rng.uniform(0.4, 1.6))), col, -1)
        else:
            cv2.circle(img, (x, y), s // 2, col, -1)
    return np.clip(img, 0, 255).astype(np.uint8)


def paste_face(
    img: np.ndarray,
    identity: int,
    cx: float,
    cy: float,
    side: float,
    angle_deg: float = 0.0,
    gain: float = 1.0,
    tex_cache: dict | None = None,
):
    """Paste identity's patch as a `side`-px square centred on (cx, cy).

    Returns (bbox xyxy float32[4], kps float32[5,2]) in image coordinates.
    """
    if tex_cache is not None and identity in tex_cache:
        tex = tex_cache[identity]
    else:
        tex = identity_texture(identity)
        if tex_cache is not None:
            tex_cache[identity] = tex
    ts = tex.shape[0]
    s = side / ts
    a = np.deg2rad(angle_deg)
    ca, sa = np.cos(a) * s, np.sin(a) * s
    # texture -> image affine
    M = np.array(
        [[ca, -sa, cx - (ca * ts / 2 - sa * ts / 2)], [sa, ca, cy - (sa * ts / 2 + ca * ts / 2)]],
        dtype=np.float64,
    )
    h, w = img.shape[:2]
    src = tex.astype(np.float32) * gain
    warped = cv2.warpAffine(src, M, (w, h), flags=cv2.INTER_AREA if s < 1 else cv2.INTER_LINEAR,
                            borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    mask = cv2.warpAffine(np.ones((ts, ts), np.float32), M, (w, h), flags=cv2.INTER_LINEAR,
                          borderMode=cv2.BORDER_CONSTANT, borderValue=0)[..., None]
    out = img.astype(np.float32) * (1 - mask) + warped * mask
    img[...] = np.clip(out, 0, 255).astype(np.uint8)
    corners = np.array([[0, 0], [ts, 0], [ts, ts], [0, ts]], dtype=np.float64)
    cw = corners @ M[:, :2].T + M[:, 2]
    bbox = np.array([cw[:, 0].min(), cw[:, 1].min(), cw[:, 0].max(), cw[:, 1].max()], dtype=np.float32)
    kps = ((_LM_FRAC.astype(np.float64) * ts) @ M[:, :2].T + M[:, 2]).astype(np.float32)
    return bbox, kps


def reference_image(identity: int, size: int = 512, seed: int = 0) -> np.ndarray:
    """A `size`x`size` reference photo: one large face of `identity` on a background."""
    rng = np.random.default_rng(1_000_000 + seed)
    img = background(rng, size, size, clutter=3)
    paste_face(img, identity, size / 2, size / 2, size * 0.45)
    return img


class ClipSpec:
    """Deterministic description of a synthetic clip (frames are rendered on demand)."""

    def __init__(self, width: int, height: int, n_frames: int, seed: int, target: int = 1,
                 others: tuple = (2, 3, 4), faces_per_frame: int = 1, face_px=(48, 120),
                 target_segments=None, fps: float = 24.0, crowd: int = 0, distractor_prob: float = 0.4):
        self.width, self.height, self.n_frames, self.seed = width, height, n_frames, seed
        self.target, self.others = target, tuple(others)
        self.faces_per_frame = faces_per_frame
        self.face_px = face_px
        self.fps = fps
        self.crowd = crowd
        self.distractor_prob = float(distractor_prob)
        if target_segments is None:
            # target visible in two stretches, absent elsewhere
            a, b = int(0.15 * n_frames), int(0.40 * n_frames)
            c, d = int(0.62 * n_frames), int(0.85 * n_frames)
            target_segments = [(a, b), (c, d)]
        self.target_segments = list(target_segments)
        self._tex = {}

    def target_visible(self, i: int) -> bool:
        return any(s <= i <= e for s, e in self.target_segments)

    def frame(self, i: int) -> np.ndarray:
        frame, _ = self.frame_with_truth(i)
        return frame

    def frame_with_truth(self, i: int):
        """Returns (BGR uint8 [H,W,3], list of (identity, bbox, kps))."""
        W, H = self.width, self.height
        # background changes slowly: one background per 48-frame block
        brng = np.random.default_rng(self.seed * 1_000_003 + (i // 48))
        img = background(brng, H, W, clutter=5)
        rng = np.random.default_rng(self.seed * 7_000_001 + i)
        truth = []
        scale = min(W, H) / 360.0
        lo, hi = self.face_px[0] * scale, self.face_px[1] * scale
        if self.crowd > 0:
            # grid of non-overlapping faces with jitter (config C4: ~64 faces)
            cols = int(np.ceil(np.sqrt(self.crowd * W / H)))
            rows = int(np.ceil(self.crowd / cols))
            cw, ch = W / cols, H / rows
            k = 0
            for r in range(rows):
                for c in range(cols):
                    if k >= self.crowd:
                        break
                    ident = self.target if (k == 0 and self.target_visible(i)) else 100 + (k % 997)
                    side = float(rng.uniform(0.45, 0.8) * min(cw, ch))
                    cx = (c + 0.5) * cw + float(rng.uniform(-0.08, 0.08)) * cw
                    cy = (r + 0.5) * ch + float(rng.uniform(-0.08, 0.08)) * ch
                    bb, kp = paste_face(img, ident, cx, cy, side, float(rng.uniform(-6, 6)),
                                        float(rng.uniform(0.9, 1.1)), self._tex)
                    truth.append((ident, bb, kp))
                    k += 1
            return img, truth
        idents = []
        if self.target_visible(i):
            idents.append(self.target)
        # a distractor identity in roughly 40% of frames (blocks of 24 frames)
        drng = np.random.default_rng(self.seed * 13 + (i // 24))
        if self.others and drng.random() < self.distractor_prob:
            idents.append(int(self.others[int(drng.integers(0, len(self.others)))]))
        slots = []
        for ident in idents[: max(1, self.faces_per_frame + 1)]:
            # smooth motion: position is a slow sinusoid + small jitter
            ph = (ident * 0.37) % 1.0
            side = float(lo + (hi - lo) * (0.5 + 0.5 * np.sin(2 * np.pi * (i / 180.0 + ph))))
            for _try in range(8):
                cx = W * (0.5 + 0.32 * np.sin(2 * np.pi * (i / 240.0 + ph + 0.11 * _try))) + float(rng.uniform(-2, 2))
                cy = H * (0.5 + 0.25 * np.cos(2 * np.pi * (i / 200.0 + 2 * ph + 0.07 * _try))) + float(rng.uniform(-2, 2))
                ok = all(abs(cx - x) > (side + s) * 0.6 or abs(cy - y) > (side + s) * 0.6 for x, y, s in slots)
                if ok:
                    break
            cx = float(np.clip(cx, side * 0.6, W - side * 0.6))
            cy = float(np.clip(cy, side * 0.6, H - side * 0.6))
            slots.append((cx, cy, side))
            bb, kp = paste_face(img, ident, cx, cy, side, float(rng.uniform(-5, 5)),
                                float(rng.uniform(0.92, 1.08)), self._tex)
            truth.append((ident, bb, kp))
        return img, truth
