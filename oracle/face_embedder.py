"""ORACLE (test infrastructure): CPU restatement of the reference FaceEmbedder SCRFD+ArcFace path.

Follows person_capture/face_embedder.py:
  extract / _extract_with_scrfd            :1663-1669, :2095-2103
  _extract_with_scrfd_raw (policy, passes) :2163-2482
  _iou, best_face                          :2484-2508
  _face_quality, _ARC_DST, preprocess      :1274-1288
  _arcface_encode (generic branch)         :1290-1308, :1369-1389
  _canon_5pts, _align_by_5pts              :1430-1473
  _upright_by_eye_roll                     :1571-1647
  state + knobs                            :473-500, :1224-1272
The networks are the folded torch-CPU executors of oracle/models.py (stand-in for ONNX
Runtime CPU, which is not installed); every image operation is the real cv2 call the
reference makes.  Pinned against the reference itself: tests/golden/reference_golden.npz holds what the
unmodified `FaceEmbedder` (sessions substituted, tests/golden/ref_harness.py) returned for 18 alignment unit cases and a
53-call extract script (228 recorded SCRFD passes); tests/test_cpu_reference_golden.py replays the detector and demands the same
passes (input image CRC, size, threshold, order), chips, boxes, qualities, features and streak / rotation state from this module.

`id(self) & 7` in the adaptive-rotation period test (face_embedder.py:2338) is process
dependent in the reference; the oracle exposes it as `rot_phase` (default 0).
"""
from __future__ import annotations

import math
from typing import List, Optional

import cv2
import numpy as np

ARC_DST = np.array([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366],
                    [41.5493, 92.3655], [70.7299, 92.2041]], dtype=np.float32)


def round32(x: int) -> int:  # face_embedder.py:86-87
    return ((int(x) + 31) // 32) * 32


def iou_int(a, b) -> float:  # face_embedder.py:2484-2494
    iw = max(0, min(a[2], b[2]) - max(a[0], b[0]))
    ih = max(0, min(a[3], b[3]) - max(a[1], b[1]))
    inter = iw * ih
    ua = max(0, a[2] - a[0]) * max(0, a[3] - a[1]) + max(0, b[2] - b[0]) * max(0, b[3] - b[1]) - inter
    return inter / ua if ua > 0 else 0.0


def face_quality(chip_bgr: np.ndarray) -> float:  # face_embedder.py:1274-1276
    g = cv2.cvtColor(chip_bgr, cv2.COLOR_BGR2GRAY)
    return float(cv2.Laplacian(g, cv2.CV_64F).var())


def canon_5pts(pts) -> Optional[np.ndarray]:  # face_embedder.py:1430-1463
    if pts is None or pts.shape != (5, 2):
        return None
    pts = np.asarray(pts, dtype=np.float32)
    if not np.isfinite(pts).all():
        return None
    by_y = np.argsort(pts[:, 1])
    eyes = pts[by_y[:2]]
    nose = pts[by_y[2]]
    mouth = pts[by_y[3:]]
    le, re = eyes[np.argsort(eyes[:, 0])]
    lm, rm = mouth[np.argsort(mouth[:, 0])]
    if not (le[0] < re[0] and lm[0] < rm[0]):
        return None
    if not (nose[1] > max(le[1], re[1]) and nose[1] < min(lm[1], rm[1])):
        return None
    return np.stack([le, re, nose, lm, rm], axis=0)


def resize_112(img: np.ndarray) -> np.ndarray:
    h, w = img.shape[:2]
    interp = cv2.INTER_AREA if max(h, w) > 112 else cv2.INTER_LINEAR
    return cv2.resize(img, (112, 112), interpolation=interp)


def align_by_5pts(bgr: np.ndarray, pts5: np.ndarray) -> np.ndarray:  # face_embedder.py:1465-1473
    M, _ = cv2.estimateAffinePartial2D(pts5.astype(np.float32), ARC_DST, method=cv2.LMEDS)
    if M is None:
        M, _ = cv2.estimateAffinePartial2D(pts5[:3], ARC_DST[:3], method=cv2.LMEDS)
    if M is None:
        return resize_112(bgr)
    return cv2.warpAffine(bgr, M, (112, 112), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)


def upright_by_eye_roll(bgr: np.ndarray, pts5: np.ndarray) -> np.ndarray:  # face_embedder.py:1571-1647
    if bgr is None or bgr.size == 0:
        return bgr
    h, w = bgr.shape[:2]
    if h == 0 or w == 0:
        return bgr
    pts = np.asarray(pts5, dtype=np.float32)
    if pts.ndim != 2 or pts.shape[0] < 5 or pts.shape[1] < 2 or not np.isfinite(pts[:5, :2]).all():
        return resize_112(bgr)
    c = pts[:5, :2].copy()
    c[:, 0] = np.clip(c[:, 0], 0.0, max(0, w - 1))
    c[:, 1] = np.clip(c[:, 1], 0.0, max(0, h - 1))
    vec = c[1] - c[0]
    if float(np.hypot(vec[0], vec[1])) < 1e-3:
        vec = c[4] - c[3]
        if float(np.hypot(vec[0], vec[1])) < 1e-3:
            return resize_112(bgr)
    angle = math.degrees(math.atan2(float(vec[1]), float(vec[0])))
    if angle < -90.0:
        angle += 180.0
    elif angle > 90.0:
        angle -= 180.0
    if abs(angle) < 8.0:
        return resize_112(bgr)
    if angle > 80.0:
        angle = 90.0
    elif angle < -80.0:
        angle = -90.0
    side = max(h, w)
    scale = 1.0 if side <= 256 else 256.0 / float(side)
    M = cv2.getRotationMatrix2D((w / 2.0, h / 2.0), -angle, scale)
    rotated = cv2.warpAffine(bgr, M, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
    pts_h = np.hstack([pts[:5, :2], np.ones((5, 1), dtype=np.float32)])
    pts_rot = (M @ pts_h.T).T.astype(np.float32)
    canon = canon_5pts(pts_rot)
    if canon is not None:
        return align_by_5pts(rotated, canon.astype(np.float32))
    return resize_112(rotated)


def arcface_preprocess(bgr: np.ndarray) -> np.ndarray:  # face_embedder.py:1281-1288
    rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
    if rgb.shape[:2] != (112, 112):
        interp = cv2.INTER_AREA if (rgb.shape[0] > 112 or rgb.shape[1] > 112) else cv2.INTER_LINEAR
        rgb = cv2.resize(rgb, (112, 112), interpolation=interp)
    arr = rgb.astype(np.float32) / 127.5 - 1.0
    return np.transpose(arr, (2, 0, 1))


def _unrotate(xr, yr, deg, W0, H0):  # face_embedder.py:2170-2175
    if deg == 90:
        return yr, H0 - 1 - xr
    if deg == 180:
        return W0 - 1 - xr, H0 - 1 - yr
    if deg == 270:
        return W0 - 1 - yr, xr
    return xr, yr


_ROT = {90: cv2.ROTATE_90_CLOCKWISE, 180: cv2.ROTATE_180, 270: cv2.ROTATE_90_COUNTERCLOCKWISE}


class FaceEmbedderOracle:
    """Same public surface as the reference class for the scrfd + arcface configuration."""

    def __init__(self, scrfd, arcface, conf: float = 0.30, rot_phase: int = 0):
        self.scrfd = scrfd          # SCRFDOracle
        self.arc = arcface          # object with run(float32[n,3,112,112]) -> float32[n,512]
        self.conf = float(conf)
        self.backend = "scrfd"
        self.use_arcface = True
        self.scrfd_tta_scales = (0.75, 0.60)
        self.scrfd_probe_conf_cap = 0.20
        self.scrfd_edge_pad_frac = 0.06
        self.scrfd_min_box_px = 8
        self._fast_prescan = False
        self._prescan_rr = 0
        self._prescan_rr_mode = "rr"
        self._prescan_escalate = False
        self._probe_conf = 0.03
        self._high_90 = 1536
        self._high_180 = 1280
        self._prescan_period = 3
        self._prescan_probe_imgsz = 384
        self._prescan_no_upscale_det = True
        self._heavy_cap = 2048
        self._frame_idx = 0
        self._no_face_streak = 0
        self._last_face_idx = -10 ** 9
        self._rot_cycle = 0
        self.rot_adaptive = True
        self.rot_every_n = 12
        self.rot_after_hit_frames = 8
        self.fast_no_face_imgsz = 512
        self.rot_phase = int(rot_phase) & 7
        self.trace: List[dict] = []   # per-call record of SCRFD passes (for tests)
        self.keep_trace = False

    # ---- knobs (face_embedder.py:1224-1272) ----
    def set_prescan_fast(self, enable: bool, *, mode: str = "rr") -> None:
        self._fast_prescan = bool(enable)
        self._prescan_rr_mode = str(mode)
        if enable:
            self._prescan_rr = 0

    def set_prescan_hint(self, *, escalate: bool = False) -> None:
        self._prescan_escalate = bool(escalate)

    def configure_rotation_strategy(self, *, adaptive=None, every_n=None, after_hit_frames=None,
                                    fast_no_face_imgsz=None) -> None:
        if adaptive is not None:
            self.rot_adaptive = bool(adaptive)
        if every_n is not None:
            self.rot_every_n = max(1, int(every_n))
        if after_hit_frames is not None:
            self.rot_after_hit_frames = max(0, int(after_hit_frames))
        if fast_no_face_imgsz is not None:
            self.fast_no_face_imgsz = max(0, int(fast_no_face_imgsz))
        self._rot_cycle = 0

    @staticmethod
    def best_face(faces):  # face_embedder.py:2504-2508
        if not faces:
            return None
        return max(faces, key=lambda f: (f["quality"], (f["bbox"][2] - f["bbox"][0]) * (f["bbox"][3] - f["bbox"][1])))

    # ---- detection pass ----
    def _detect_once(self, img, size, conf):
        self.scrfd.det_thresh = float(conf)
        det, kps = self.scrfd.detect(img, input_size=(size, size))
        if self.keep_trace:
            self.trace[-1]["passes"].append(dict(shape=img.shape[:2], size=size, conf=float(conf), n=len(det)))
        return det, kps

    def extract(self, bgr_img, *, imgsz: Optional[int] = None):
        if bgr_img is None or bgr_img.size == 0:
            return []
        self._frame_idx += 1
        if self.keep_trace:
            self.trace.append(dict(passes=[]))
        return self._extract_raw(bgr_img, imgsz)

    def _extract_raw(self, bgr_img, imgsz):
        H0, W0 = bgr_img.shape[:2]
        dyn = int(imgsz) if (imgsz is not None and imgsz > 0) else 640
        if self._no_face_streak >= 3:
            dyn = min(dyn, self.fast_no_face_imgsz)
        if self._fast_prescan:
            dyn = min(dyn, int(self._prescan_probe_imgsz))
            if bool(self._prescan_no_upscale_det):
                dyn = min(dyn, max(320, (max(H0, W0) // 32) * 32))
        dyn = round32(max(320, dyn))
        L = max(H0, W0)
        heavy_cap = max(int(self._heavy_cap), dyn)
        heavy90 = min(round32(max(dyn, int(0.75 * L))), heavy_cap)
        heavy180 = min(round32(max(dyn, int(0.67 * L))), heavy_cap)

        dets = []

        def accumulate(bb, kp, deg):
            x1, y1, x2, y2 = [int(v) for v in bb[:4]]
            ax, ay = _unrotate(x1, y1, deg, W0, H0)
            bx, by = _unrotate(x2, y2, deg, W0, H0)
            xa1, ya1 = min(ax, bx), min(ay, by)
            xa2, ya2 = max(ax, bx), max(ay, by)
            xa1 = max(0, min(W0 - 1, xa1))
            ya1 = max(0, min(H0 - 1, ya1))
            xa2 = max(xa1 + 1, min(W0, xa2))
            ya2 = max(ya1 + 1, min(H0, ya2))
            if xa2 - xa1 <= 2 or ya2 - ya1 <= 2:
                return
            pts = None
            if kp is not None:
                raw = np.asarray(kp, dtype=np.float32).reshape(-1, 2)
                mapped = []
                for px, py in raw:
                    ox, oy = _unrotate(float(px), float(py), deg, W0, H0)
                    mapped.append([float(ox - xa1), float(oy - ya1)])
                pts = np.asarray(mapped[:5], dtype=np.float32) if len(mapped) >= 5 else None
            dets.append(((xa1, ya1, xa2, ya2), pts, float(bb[4]) if len(bb) > 4 else 1.0))

        def collect(bbs, kpss, deg, fix=None):
            if bbs is None or len(bbs) == 0:
                return
            for i, bb in enumerate(bbs):
                kp = None if (kpss is None or i >= len(kpss)) else kpss[i]
                if fix is not None:
                    bb, kp = fix(np.asarray(bb).copy(), None if kp is None else np.asarray(kp).copy())
                accumulate(bb, kp, deg)

        # pass 1: upright
        bbs, kpss = self._detect_once(bgr_img, dyn, self.conf)
        collect(bbs, kpss, 0)

        # normal mode only: scale TTA then replicate-pad probe (face_embedder.py:2251-2315)
        if not dets and not self._fast_prescan:
            scales = tuple(self.scrfd_tta_scales) + ((1.25,) if max(W0, H0) <= 1920 else ())
            probe_conf = min(float(self.conf), float(self.scrfd_probe_conf_cap))
            for s in scales:
                if s == 1.0:
                    continue
                interp = cv2.INTER_AREA if s < 1.0 else cv2.INTER_LINEAR
                img_s = cv2.resize(bgr_img, None, fx=s, fy=s, interpolation=interp)
                dyn_s = round32(min(self._heavy_cap, max(320, int(dyn * s))))
                bb_s, kp_s = self._detect_once(img_s, dyn_s, probe_conf)
                inv = 1.0 / s

                def fix_scale(bb, kp, inv=inv):
                    bb[:4] = np.asarray(bb[:4], dtype=np.float32) * inv
                    if kp is not None:
                        kp = np.asarray(kp, dtype=np.float32) * inv
                    return bb, kp

                collect(bb_s, kp_s, 0, fix_scale)
                if dets:
                    break
            if not dets:
                pad = int(round(min(64, float(self.scrfd_edge_pad_frac) * max(W0, H0))))
                if pad > 0:
                    img_p = cv2.copyMakeBorder(bgr_img, pad, pad, pad, pad, cv2.BORDER_REPLICATE)
                    bb_p, kp_p = self._detect_once(img_p, dyn, probe_conf)

                    def fix_pad(bb, kp, pad=pad):
                        bb[:4] -= np.array([pad, pad, pad, pad], dtype=np.float32)
                        bb[0] = max(0.0, min(float(W0 - 1), float(bb[0])))
                        bb[1] = max(0.0, min(float(H0 - 1), float(bb[1])))
                        bb[2] = max(bb[0] + 1.0, min(float(W0), float(bb[2])))
                        bb[3] = max(bb[1] + 1.0, min(float(H0), float(bb[3])))
                        if kp is not None:
                            kp = np.asarray(kp, dtype=np.float32)
                            kp[..., 0] = np.clip(kp[..., 0] - pad, 0, W0 - 1)
                            kp[..., 1] = np.clip(kp[..., 1] - pad, 0, H0 - 1)
                        return bb, kp

                    collect(bb_p, kp_p, 0, fix_pad)

        mp = int(self.scrfd_min_box_px)
        dets = [d for d in dets if (d[0][2] - d[0][0] >= mp and d[0][3] - d[0][1] >= mp)]

        # rotation policy (face_embedder.py:2330-2360)
        if not dets:
            need_rot = False
            self._no_face_streak += 1
            if self.rot_adaptive:
                if (self._frame_idx - self._last_face_idx) <= self.rot_after_hit_frames:
                    need_rot = True
                elif ((self._frame_idx + self.rot_phase) % self.rot_every_n) == 0:
                    need_rot = True
            else:
                need_rot = True
        else:
            need_rot = False
            self._no_face_streak = 0
            self._last_face_idx = self._frame_idx
            self._rot_cycle = 0
        if self._fast_prescan:
            if dets:
                need_rot = False
            else:
                period = max(1, int(self._prescan_period))
                need_rot = need_rot or self._prescan_escalate or (((self._frame_idx + self._prescan_rr) % period) == 0)
        if self._fast_prescan and not dets and not need_rot:
            return []

        if not dets and need_rot:
            self._rot_cycle += 1
            if self._fast_prescan:
                if self._prescan_rr_mode == "rr":
                    rot_seq = ((90, 270)[self._prescan_rr % 2],)
                    self._prescan_rr += 1
                else:
                    rot_seq = (90, 270)
            else:
                rot_seq = (90, 270, 180)
            for deg in rot_seq:
                rimg_probe = cv2.rotate(bgr_img, _ROT[deg])
                probe_conf = max(0.02, float(self._probe_conf))
                probe_dyn = round32(max(320, min(dyn, int(self._prescan_probe_imgsz))))
                pb, _ = self._detect_once(rimg_probe, probe_dyn, probe_conf)
                hits = len(pb) if pb is not None else 0
                do_heavy = (hits > 0) or (self._fast_prescan and self._prescan_escalate) or (not self._fast_prescan)
                if self._fast_prescan and hits == 0:
                    continue
                pad = 24
                rimg = cv2.copyMakeBorder(rimg_probe, pad, pad, pad, pad, cv2.BORDER_REPLICATE)
                if self._fast_prescan:
                    heavy = heavy180 if deg == 180 else heavy90
                    override = self._high_180 if deg == 180 else self._high_90
                    if override and override > 0:
                        heavy = max(heavy, round32(int(override)))
                    heavy = min(heavy, int(self._heavy_cap))
                    det_sizes = [heavy] if do_heavy else [dyn]
                else:
                    sizes = []
                    for base in (max(dyn, 1280), max(dyn, 1536)):
                        base = round32(base)
                        if base not in sizes:
                            sizes.append(base)
                    det_sizes = sizes if do_heavy else [dyn]
                conf_deg = max(0.10, float(self.conf) * (0.8 if deg in (90, 270) else 0.6))
                rb = rk = None
                for ds in det_sizes:
                    rb, rk = self._detect_once(rimg, ds, conf_deg)
                    if rb is not None and len(rb) > 0:
                        break
                    rb = rk = None
                if rb is None:
                    continue

                def fix_unpad(bb, kp, pad=pad):
                    bb[:4] -= np.array([pad, pad, pad, pad], dtype=bb.dtype)
                    if kp is not None:
                        kp[..., 0] -= pad
                        kp[..., 1] -= pad
                    return bb, kp

                collect(rb, rk, deg, fix_unpad)
                if dets:
                    break
        if not dets:
            return []

        # cross-pass suppression (face_embedder.py:2439-2443)
        dets.sort(key=lambda t: (t[2], (t[0][2] - t[0][0]) * (t[0][3] - t[0][1])), reverse=True)
        kept = []
        for box, pts, sc in dets:
            if all(iou_int(box, k[0]) < 0.45 for k in kept):
                kept.append((box, pts, sc))

        metas, chips = [], []
        for (x1, y1, x2, y2), kps, _sc in kept:
            xi1 = max(0, min(W0 - 1, int(round(x1))))
            yi1 = max(0, min(H0 - 1, int(round(y1))))
            xi2 = max(xi1 + 1, min(W0, int(round(x2))))
            yi2 = max(yi1 + 1, min(H0, int(round(y2))))
            crop = bgr_img[yi1:yi2, xi1:xi2]
            chip = None
            if kps is not None:
                pts = np.asarray(kps, dtype=np.float32)
                canon = canon_5pts(pts)
                chip = align_by_5pts(crop, canon) if canon is not None else upright_by_eye_roll(crop, pts)
            if chip is None:
                chip = resize_112(crop)
            metas.append((xi1, yi1, xi2, yi2, face_quality(chip)))
            chips.append(chip)
        self.last_chips = chips
        feats = self.arcface_encode(chips)
        out = [dict(bbox=np.array(m[:4], dtype=np.int32), feat=feats[i], quality=float(m[4]))
               for i, m in enumerate(metas)]
        out.sort(key=lambda f: (f["quality"], (f["bbox"][2] - f["bbox"][0]) * (f["bbox"][3] - f["bbox"][1])), reverse=True)
        return out

    def arcface_encode(self, chips):  # face_embedder.py:1290-1389 (generic branch)
        if not chips:
            return []
        m = len(chips)
        do_flip = (not self._fast_prescan) or self._prescan_escalate
        pairs = list(chips)
        if do_flip:
            pairs.extend(cv2.flip(b, 1) for b in chips)
        X = np.empty((len(pairs), 3, 112, 112), dtype=np.float32)
        for i, b in enumerate(pairs):
            X[i] = arcface_preprocess(b)
        feats = np.asarray(self.arc.run(np.ascontiguousarray(X)), dtype=np.float32)
        f = feats[:m].copy()
        if do_flip:
            f += feats[m:2 * m]
        norms = np.linalg.norm(f, axis=1, keepdims=True).astype(np.float32, copy=False)
        np.maximum(norms, 1e-6, out=norms)
        f /= norms
        return f.astype(np.float32, copy=False)
