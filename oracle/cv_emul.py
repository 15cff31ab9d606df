"""ORACLE (test infrastructure): numpy restatement of the OpenCV uint8 kernels the path relies on.

Third-party dependency: opencv-python-headless==4.9.0.80 (reference requirements.txt:10; this
image has 4.13.0).  Call sites in the reference: cv2.resize (gui_app.py:1505-1507,
face_embedder.py:1285, 1472, 2264 and inside InsightFace SCRFD.detect), cv2.warpAffine
(face_embedder.py:1473, 1630), cv2.estimateAffinePartial2D(LMEDS) (face_embedder.py:1466),
cv2.cvtColor/Laplacian (face_embedder.py:1275-1276).  The CUDA kernels (csrc/pcb_cvmath.h)
implement exactly these formulas; tests/test_cv_emul.py pins this file against the real cv2
calls, so cv2 itself is the golden reference ("parity pinned against cv2 4.13").
Slow (numpy loops over small dims); only for tests.
"""
from __future__ import annotations

import math

import numpy as np


def _rint(x):
    return np.rint(x)


# ---- INTER_LINEAR uint8 (resize.cpp: resizeGeneric_ with HResizeLinear/VResizeLinear fixed point)
def linear_coeffs(src: int, dst: int, horizontal: bool, area_mode: bool = False):
    inv = float(dst) / float(src)
    scale = 1.0 / inv
    idx = np.zeros(dst, np.int64)
    a0 = np.zeros(dst, np.int64)
    a1 = np.zeros(dst, np.int64)
    for d in range(dst):
        if area_mode:
            s = int(math.floor(d * scale))
            f = np.float32((d + 1) - (s + 1) * inv)
            f = np.float32(0.0) if f <= 0 else np.float32(f - np.float32(math.floor(f)))
        else:
            f = np.float32((d + 0.5) * scale - 0.5)
            s = int(math.floor(f))
            f = np.float32(f - np.float32(s))
        if horizontal:
            if s < 0:
                s, f = 0, np.float32(0.0)
            if s >= src - 1:
                s, f = src - 1, np.float32(0.0)
        idx[d] = s
        a0[d] = int(_rint(np.float32(np.float32(1.0) - f) * np.float32(2048.0)))
        a1[d] = int(_rint(np.float32(f) * np.float32(2048.0)))
    return idx, a0, a1


def resize_linear_u8(img: np.ndarray, dw: int, dh: int, area_mode: bool = False) -> np.ndarray:
    h, w = img.shape[:2]
    if not area_mode and w == 2 * dw and h == 2 * dh:
        return resize_area_u8(img, dw, dh)  # cv::resize: INTER_LINEAR with exact 2x == INTER_AREA
    xi, xa0, xa1 = linear_coeffs(w, dw, True, area_mode)
    yi, ya0, ya1 = linear_coeffs(h, dh, False, area_mode)
    S = img.astype(np.int64)
    x1 = np.minimum(xi + 1, w - 1)
    H = S[:, xi] * xa0[None, :, None] + S[:, x1] * xa1[None, :, None]  # [h, dw, c]
    y0 = np.clip(yi, 0, h - 1)
    y1 = np.clip(yi + 1, 0, h - 1)
    r0 = (ya0[:, None, None] * (H[y0] >> 4)) >> 16
    r1 = (ya1[:, None, None] * (H[y1] >> 4)) >> 16
    return np.clip((r0 + r1 + 2) >> 2, 0, 255).astype(np.uint8)


# ---- INTER_AREA uint8 (resize.cpp: resizeAreaFast_ / resizeArea_)
def area_table(src: int, dst: int):
    scale = 1.0 / (float(dst) / float(src))
    tab = []
    for d in range(dst):
        fs1 = d * scale
        fs2 = fs1 + scale
        cell = min(scale, src - fs1)
        s1 = int(math.ceil(fs1))
        s2 = int(math.floor(fs2))
        s2 = min(s2, src - 1)
        s1 = min(s1, s2)
        if s1 - fs1 > 1e-3:
            tab.append((d, s1 - 1, np.float32((s1 - fs1) / cell)))
        for s in range(s1, s2):
            tab.append((d, s, np.float32(1.0 / cell)))
        if fs2 - s2 > 1e-3:
            tab.append((d, s2, np.float32(min(min(fs2 - s2, 1.0), cell) / cell)))
    return tab


def resize_area_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    h, w = img.shape[:2]
    sx = 1.0 / (float(dw) / float(w))
    sy = 1.0 / (float(dh) / float(h))
    if sx < 1 or sy < 1:
        return resize_linear_u8(img, dw, dh, area_mode=True)
    isx, isy = int(sx), int(sy)   # saturate_cast<int>(double) rounds; equal for exact integers
    fast = abs(sx - round(sx)) < np.finfo(np.float64).eps and abs(sy - round(sy)) < np.finfo(np.float64).eps
    if fast:
        isx, isy = int(round(sx)), int(round(sy))
        S = img[: dh * isy, : dw * isx].astype(np.int64).reshape(dh, isy, dw, isx, -1).sum(axis=(1, 3))
        if isx == 2 and isy == 2:
            return ((S + 2) >> 2).astype(np.uint8)
        sc = np.float32(1.0 / (isx * isy))
        return np.clip(_rint(S.astype(np.float32) * sc), 0, 255).astype(np.uint8)
    xt = area_table(w, dw)
    yt = area_table(h, dh)
    c = img.shape[2]
    out = np.zeros((dh, dw, c), np.uint8)
    S = img.astype(np.float32)
    summ = np.zeros((dw, c), np.float32)
    prev = yt[0][0]
    for (dy, sy_, beta) in yt:
        buf = np.zeros((dw, c), np.float32)
        row = S[sy_]
        for (dx, sx_, alpha) in xt:
            buf[dx] = buf[dx] + row[sx_] * alpha
        if dy != prev:
            out[prev] = np.clip(_rint(summ), 0, 255).astype(np.uint8)
            summ = beta * buf
            prev = dy
        else:
            summ = summ + beta * buf
    out[prev] = np.clip(_rint(summ), 0, 255).astype(np.uint8)
    return out


# ---- warpAffine INTER_LINEAR + BORDER_REFLECT uint8 (imgwarp.cpp WarpAffineInvoker + remapBilinear)
def _reflect(p, n):
    if n == 1:
        return np.zeros_like(p)
    p = p.copy()
    for _ in range(64):
        bad = (p < 0) | (p >= n)
        if not bad.any():
            break
        p = np.where(p < 0, -p - 1, p)
        p = np.where(p >= n, 2 * n - 1 - p, p)
    return p


def invert_affine_cv(M):
    M = np.array(M, dtype=np.float64).reshape(2, 3).copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11 = M[1, 1] * D
    A22 = M[0, 0] * D
    m00, m01, m10, m11 = A11, M[0, 1] * (-D), M[1, 0] * (-D), A22
    b1 = -m00 * M[0, 2] - m01 * M[1, 2]
    b2 = -m10 * M[0, 2] - m11 * M[1, 2]
    return np.array([[m00, m01, b1], [m10, m11, b2]], dtype=np.float64)


def warp_affine_u8(src: np.ndarray, M, dw: int, dh: int) -> np.ndarray:
    h, w = src.shape[:2]
    iM = invert_affine_cv(M)
    xs = np.arange(dw, dtype=np.float64)
    adelta = _rint(iM[0, 0] * xs * 1024.0).astype(np.int64)
    bdelta = _rint(iM[1, 0] * xs * 1024.0).astype(np.int64)
    out = np.zeros((dh, dw, src.shape[2]), np.uint8)
    S = src.astype(np.int64)
    for y in range(dh):
        X0 = int(_rint((iM[0, 1] * y + iM[0, 2]) * 1024.0)) + 16
        Y0 = int(_rint((iM[1, 1] * y + iM[1, 2]) * 1024.0)) + 16
        X = (X0 + adelta) >> 5
        Y = (Y0 + bdelta) >> 5
        sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)
        fx, fy = (X & 31).astype(np.float32) / np.float32(32), (Y & 31).astype(np.float32) / np.float32(32)
        w00 = _rint((1 - fy) * (1 - fx) * np.float32(32768)).astype(np.int64)
        w01 = _rint((1 - fy) * fx * np.float32(32768)).astype(np.int64)
        w10 = _rint(fy * (1 - fx) * np.float32(32768)).astype(np.int64)
        w11 = _rint(fy * fx * np.float32(32768)).astype(np.int64)
        x0, x1 = _reflect(sx, w), _reflect(sx + 1, w)
        y0, y1 = _reflect(sy, h), _reflect(sy + 1, h)
        v = (S[y0, x0] * w00[:, None] + S[y0, x1] * w01[:, None] + S[y1, x0] * w10[:, None] + S[y1, x1] * w11[:, None])
        out[y] = np.clip((v + 16384) >> 15, 0, 255).astype(np.uint8)
    return out


# ---- BGR2GRAY + Laplacian variance (face_embedder.py:1274-1276)
def gray_u8(bgr: np.ndarray) -> np.ndarray:
    b = bgr.astype(np.int64)
    # cv2 4.13 color_rgb: BY15/GY15/RY15 with a 15-bit shift (measured: 0 mismatches; the 14-bit
    # set 1868/9617/4899 of older docs mismatches on ~0.3% of pixels)
    return ((b[..., 0] * 3735 + b[..., 1] * 19235 + b[..., 2] * 9798 + 16384) >> 15).astype(np.uint8)


def laplacian_var(gray: np.ndarray) -> float:
    g = np.pad(gray.astype(np.int64), 1, mode="reflect")  # REFLECT_101
    lap = g[:-2, 1:-1] + g[2:, 1:-1] + g[1:-1, :-2] + g[1:-1, 2:] - 4 * g[1:-1, 1:-1]
    n = lap.size
    s1 = int(lap.sum())
    s2 = int((lap * lap).sum())
    return (n * s2 - s1 * s1) / float(n * n)


# ---- estimateAffinePartial2D(LMEDS) (ptsetreg.cpp: LMeDSPointSetRegistrator + AffinePartial2D callbacks)
class CvRNG:
    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else a + self.next() % (b - a)


def lmeds_pairs(count: int, niters: int = 13):
    """The (idx0, idx1) subsets LMedS draws for `count` points: RNG is re-seeded per call."""
    rng = CvRNG()
    pairs = []
    for _ in range(niters):
        i0 = rng.uniform(0, count)
        while True:
            i1 = rng.uniform(0, count)
            if i1 != i0:
                break
        pairs.append((i0, i1))
    return pairs


def _two_point_model(f, t):
    x1, y1, x2, y2 = float(f[0][0]), float(f[0][1]), float(f[1][0]), float(f[1][1])
    X1, Y1, X2, Y2 = float(t[0][0]), float(t[0][1]), float(t[1][0]), float(t[1][1])
    den = (x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2)
    with np.errstate(all="ignore"):
        d = np.float64(1.0) / np.float64(den)
        S0 = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2))
        S1 = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2))
        S2 = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2) - (X1 * x2 - X2 * x1) * (x1 - x2))
        S3 = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2) - (Y1 * y2 - Y2 * y1) * (y1 - y2))
    return np.array([S0, -S1, S2, S1, S0, S3], dtype=np.float64)


def _errors(model, src, dst):
    F = model.astype(np.float32)
    a = (F[0] * src[:, 0] + F[1] * src[:, 1] + F[2]) - dst[:, 0]
    b = (F[3] * src[:, 0] + F[4] * src[:, 1] + F[5]) - dst[:, 1]
    return (a * a + b * b).astype(np.float32)


def estimate_affine_partial_lmeds(src, dst):
    """-> (M float64 [2,3] or None, inlier mask).  Final fit = closed-form LS over inliers
    (cv2 reaches the same optimum with <=10 Levenberg-Marquardt steps)."""
    src = np.asarray(src, np.float32).reshape(-1, 2)
    dst = np.asarray(dst, np.float32).reshape(-1, 2)
    count = src.shape[0]
    best, best_med = None, np.inf
    for i0, i1 in lmeds_pairs(count):
        model = _two_point_model(src[[i0, i1]], dst[[i0, i1]])
        err = _errors(model, src, dst)
        med = float(np.sort(err)[count // 2]) if count % 2 else float(np.sort(err)[count // 2 - 1: count // 2 + 1].mean())
        if med < best_med:
            best_med, best = med, model
    if best is None:
        return None, None
    sigma = max(2.5 * 1.4826 * (1 + 5.0 / (count - 2)) * math.sqrt(best_med), 0.001)
    thr = np.float32(sigma * sigma)
    mask = _errors(best, src, dst) <= thr
    if mask.sum() < 2:
        return None, mask
    s = src[mask].astype(np.float64)
    d = dst[mask].astype(np.float64)
    ms, md = s.mean(0), d.mean(0)
    sc, dc = s - ms, d - md
    den = (sc * sc).sum()
    a = (sc[:, 0] * dc[:, 0] + sc[:, 1] * dc[:, 1]).sum() / den
    b = (sc[:, 0] * dc[:, 1] - sc[:, 1] * dc[:, 0]).sum() / den
    tx = md[0] - (a * ms[0] - b * ms[1])
    ty = md[1] - (b * ms[0] + a * ms[1])
    return np.array([[a, -b, tx], [b, a, ty]], dtype=np.float64), mask
