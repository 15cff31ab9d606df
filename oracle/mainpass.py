"""ORACLE (test infrastructure): CPU restatement of the identity decisions of the reference's main pass.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this package.

Restates, for the face-only pipeline (match_mode="face_only", disable_reid=True, skip_yolo_when_faceonly -- the defaults
SURVEY.md App. C describes), what person_capture/gui_app.py does to decide "is the target in this frame":
  * segment gate + seek cooldown              gui_app.py:5648-5684
  * frame_stride gate                         gui_app.py:5743-5746
  * lock-face ROI probe                       gui_app.py:5796-5855 (geometry: _expand_xyxy :4186-4199), accept :5919
  * lock-ROI miss counter                     gui_app.py:6024-6028
  * full-frame cadence probe                  gui_app.py:6030-6047, best face :6058-6071, accept :6116
  * face-only global fallback                 gui_app.py:7521-7551
  * lock-face box update on accept            gui_app.py:7501-7505 (_set_lock_face_box :4164-4177)
  * runtime bank learning (off by default)    gui_app.py:7460-7494
  * cooldown decrement                        gui_app.py:8003-8006
  * per-person-crop site (boxes given)      gui_app.py:6269-6346, 6370-6437   -> person_crop_candidates
  * frame-level arbitration + lock gate       gui_app.py:7788-7845              -> arbitrate
Crop composition (ratio choice, margins, sharpness of the composed crop), the person detector and saving are product logic
downstream of identity and out of scope: candidates carry the person box as their area and sharp = 0.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import prescan as OP


def expand_xyxy(box, pad_x: float, pad_y: float, frame_w: int, frame_h: int) -> Tuple[int, int, int, int]:
    """gui_app.py:4186-4199."""
    x1, y1, x2, y2 = [float(v) for v in box]
    ix1 = max(0, min(frame_w - 1, int(math.floor(x1 - pad_x))))
    iy1 = max(0, min(frame_h - 1, int(math.floor(y1 - pad_y))))
    ix2 = max(ix1 + 1, min(frame_w, int(math.ceil(x2 + pad_x))))
    iy2 = max(iy1 + 1, min(frame_h, int(math.ceil(y2 + pad_y))))
    return ix1, iy1, ix2, iy2


def pick_best(faces: Sequence[dict], bank, quality_min: float, use_quality_vis: bool):
    """gui_app.py:5855-5866 / 6058-6071 / 7534-7548: argmin fd over quality-passing faces, else over all faces with a feature."""
    with_feat = [f for f in faces if f.get("feat") is not None]
    if bank is not None and with_feat:
        cand = with_feat
        if use_quality_vis:
            cand = [f for f in cand if float(f.get("quality", 0.0)) >= quality_min]
        return min(cand or with_feat, key=lambda f: OP.fd_min(f["feat"], bank))
    if not faces:
        return None
    return max(faces, key=lambda f: (f["quality"], (f["bbox"][2] - f["bbox"][0]) * (f["bbox"][3] - f["bbox"][1])))


def main_pass(get_frame: Callable[[int], Optional[np.ndarray]], fps: float, total_frames: int, keep_spans: Sequence[Tuple[int, int]],
              face, ref_face_feat, cfg, log: Optional[list] = None):
    """Sequential main pass over the kept spans.  -> list of accepted hits
    dict(idx, site, fd, quality, face_box (x1,y1,x2,y2) ints in frame px)."""
    bank = None if ref_face_feat is None else np.asarray(ref_face_feat, np.float32).reshape(-1, 512)
    bank_list = [] if bank is None else [r.copy() for r in bank]
    stride = max(1, int(getattr(cfg, "frame_stride", 2)))
    face_thresh = float(cfg.face_thresh)
    quality_min = float(cfg.face_quality_min)
    use_qv = bool(getattr(cfg, "face_visible_uses_quality", True))
    imgsz = getattr(cfg, "face_fullframe_imgsz", None)
    imgsz = int(imgsz) if imgsz is not None else None
    roi_enable = bool(getattr(cfg, "lock_face_roi_enable", True))
    roi_pad = max(0.0, float(getattr(cfg, "lock_face_roi_pad", 1.25)))
    max_misses = max(0, int(getattr(cfg, "lock_face_roi_max_misses", 8)))
    cadence = int(getattr(cfg, "face_fullframe_cadence", 12))
    when_missed = bool(getattr(cfg, "face_fullframe_when_missed", True))
    learn = bool(getattr(cfg, "learn_bank_runtime", False))
    lock_box = None
    misses = 0
    cooldown = 0
    last_add = -10 ** 9
    hits: List[dict] = []
    span_i = 0
    frame_idx = 0
    spans = list(keep_spans)
    while frame_idx < total_frames:
        if spans:
            if span_i >= len(spans):
                break
            s, e = spans[span_i]
            if frame_idx < s:                      # segment jump (gui_app.py:5652-5665)
                frame_idx = s
                cooldown = int(max(2, (fps or 30) * 0.25))
                continue
            if frame_idx > e:
                span_i += 1
                if span_i >= len(spans):
                    break
                s2, _ = spans[span_i]
                if frame_idx < s2:
                    frame_idx = s2
                    cooldown = int(max(2, (fps or 30) * 0.25))
                continue
        idx = frame_idx
        if idx % stride != 0:
            frame_idx += 1
            continue
        frame = get_frame(idx)
        if frame is None:
            break
        H2, W2 = frame.shape[:2]
        cand = None
        rec = dict(idx=idx, site=None, roi=None, n_faces=0, fd=None, accept=False)
        # ---- lock-face ROI probe
        if bank is not None and roi_enable and lock_box is not None and cooldown <= 0:
            ran = False
            lx1, ly1, lx2, ly2 = [float(v) for v in lock_box]
            roi_faces = []
            rx1 = ry1 = 0
            if lx2 > 0.0 and ly2 > 0.0 and lx1 < float(W2) and ly1 < float(H2):
                fw, fh = max(1.0, lx2 - lx1), max(1.0, ly2 - ly1)
                rx1, ry1, rx2, ry2 = expand_xyxy((lx1, ly1, lx2, ly2), max(16.0, fw * roi_pad), max(16.0, fh * roi_pad), W2, H2)
                if rx2 > rx1 + 8 and ry2 > ry1 + 8:
                    ran = True
                    rec["roi"] = (rx1, ry1, rx2, ry2)
                    roi = frame[ry1:ry2, rx1:rx2]
                    if roi.size > 0:
                        roi_faces = face.extract(np.ascontiguousarray(roi), imgsz=imgsz)
            if roi_faces:
                rec["n_faces"] = len(roi_faces)
                g = pick_best(roi_faces, bank, quality_min, use_qv)
                if g is not None and g.get("feat") is not None:
                    fx1, fy1, fx2, fy2 = [float(v) for v in g["bbox"]]
                    fx1 += rx1; fx2 += rx1; fy1 += ry1; fy2 += ry1
                    fx1 = max(0.0, min(float(W2), fx1)); fy1 = max(0.0, min(float(H2), fy1))
                    fx2 = max(fx1 + 1.0, min(float(W2), fx2)); fy2 = max(fy1 + 1.0, min(float(H2), fy2))
                    fd = OP.fd_min(g["feat"], bank)
                    rec["fd"] = fd
                    if fd <= face_thresh:
                        cand = dict(site="lock_roi", fd=fd, feat=g["feat"], quality=float(g.get("quality", 0.0)), box=(fx1, fy1, fx2, fy2))
                        misses = 0
            if ran and cand is None:
                misses += 1
                if misses > max_misses:
                    lock_box = None
                    misses = 0
        # ---- full-frame probe at a coarse cadence (face-only pipeline)
        gfaces = []
        if bank is not None and when_missed and cand is None:
            if cadence <= 0 or (idx % max(1, cadence) == 0):
                gfaces = face.extract(frame, imgsz=imgsz)
        if gfaces:
            rec["n_faces"] = len(gfaces)
            g = pick_best(gfaces, bank, quality_min, use_qv)
            if g is not None and g.get("feat") is not None:
                cand = _fullframe_candidate(g, bank, face_thresh, W2, H2, "fullframe", rec)
        # ---- face-only global fallback: nothing so far -> one more full-frame extract
        if cand is None and bank is not None and when_missed:
            gfaces = face.extract(frame, imgsz=imgsz)
            rec["n_faces"] = max(rec["n_faces"], len(gfaces))
            g = pick_best(gfaces, bank, quality_min, use_qv)
            if g is not None and g.get("feat") is not None:
                cand = _fullframe_candidate(g, bank, face_thresh, W2, H2, "fallback", rec)
        if cand is not None:
            x1, y1, x2, y2 = cand["box"]
            fx1i = max(0, min(W2 - 1, int(round(x1))))
            fy1i = max(0, min(H2 - 1, int(round(y1))))
            fx2i = max(fx1i + 1, min(W2, int(round(x2))))
            fy2i = max(fy1i + 1, min(H2, int(round(y2))))
            rec.update(site=cand["site"], fd=cand["fd"], accept=True)
            hits.append(dict(idx=idx, site=cand["site"], fd=cand["fd"], quality=cand["quality"], face_box=(fx1i, fy1i, fx2i, fy2i)))
            if learn:
                ok_q = quality_min <= 0 or cand["quality"] >= quality_min
                cd = int(getattr(cfg, "prescan_add_cooldown_samples", 5)) * stride
                if cand["fd"] <= float(getattr(cfg, "prescan_fd_add", 0.22)) and ok_q and (idx - last_add) >= cd:
                    bank, action, _ = OP.bank_update(bank_list, bank, cand["feat"], cand["quality"], cfg)
                    if action in ("added", "replaced"):
                        last_add = idx
            if fx2i > fx1i and fy2i > fy1i:       # _set_lock_face_box
                lock_box = (fx1i, fy1i, fx2i, fy2i)
                misses = 0
        if log is not None:
            log.append(rec)
        frame_idx = idx + 1
        if cooldown > 0:
            cooldown -= 1
    return hits


def _fullframe_candidate(g, bank, face_thresh, W2, H2, site, rec):
    fx1, fy1, fx2, fy2 = [float(v) for v in g["bbox"]]
    fx1 = max(0.0, min(float(W2), fx1)); fy1 = max(0.0, min(float(H2), fy1))
    fx2 = max(fx1 + 1.0, min(float(W2), fx2)); fy2 = max(fy1 + 1.0, min(float(H2), fy2))
    fd = OP.fd_min(g["feat"], bank)
    rec["fd"] = fd
    if fd <= face_thresh:
        return dict(site=site, fd=fd, feat=g["feat"], quality=float(g.get("quality", 0.0)), box=(fx1, fy1, fx2, fy2))
    return None


# ---------------------------------------------------------------------------------------------------------------
# per-person-crop identity site and frame-level arbitration (SURVEY.md App. C rules 2-4)
# ---------------------------------------------------------------------------------------------------------------
def iou_xyxy(a, b) -> float:
    """gui_app.py:3484-3493."""
    ax1, ay1, ax2, ay2 = a
    bx1, by1, bx2, by2 = b
    iw, ih = max(0, min(ax2, bx2) - max(ax1, bx1)), max(0, min(ay2, by2) - max(ay1, by1))
    inter = iw * ih
    union = max(0, (ax2 - ax1) * (ay2 - ay1)) + max(0, (bx2 - bx1) * (by2 - by1)) - inter + 1e-9
    return inter / union


def person_crop_candidates(frame, boxes, face, ref_face_feat, cfg):
    """Identity kernel of the per-person site for person boxes given by the caller (gui_app.py:6269-6346, 6370-6437, face-only
    pipeline): face.extract on each crop (padded retry when empty, :6273-6293), bestf = argmin fd (:6312-6323),
    face_ok = fd <= face_thresh (:6377), accept = face_ok (:6394-6395), hard gate "face required if any face is visible"
    (:6417-6437).  -> (candidates, info); a candidate is dict(i, box, fd, score, quality, face_box, area, sharp)."""
    H2, W2 = frame.shape[:2]
    bank = None if ref_face_feat is None else np.asarray(ref_face_feat, np.float32).reshape(-1, 512)
    pad = float(getattr(cfg, "face_det_pad", 0.08))
    qmin = float(cfg.face_quality_min)
    faces_local = {}
    faces_detected = faces_passing_quality = 0
    for i, (x1, y1, x2, y2) in enumerate(boxes):
        ffaces = face.extract(np.ascontiguousarray(frame[y1:y2, x1:x2]))
        if not ffaces and pad > 0.0:
            pw, ph = int(round((x2 - x1) * pad)), int(round((y2 - y1) * pad))
            if pw > 0 or ph > 0:
                px1, py1, px2, py2 = max(0, x1 - pw), max(0, y1 - ph), min(W2, x2 + pw), min(H2, y2 + ph)
                if px2 > px1 and py2 > py1:
                    ff2 = face.extract(np.ascontiguousarray(frame[py1:py2, px1:px2]))
                    dx, dy = x1 - px1, y1 - py1
                    ffaces = []
                    for f in ff2:
                        bb = f["bbox"].copy()
                        bb[0] -= dx; bb[2] -= dx; bb[1] -= dy; bb[3] -= dy
                        ffaces.append({"bbox": bb, "feat": f.get("feat"), "quality": f.get("quality", 0.0)})
        faces_detected += len(ffaces)
        if bank is not None and ffaces:
            with_feat = [f for f in ffaces if f.get("feat") is not None]
            bestf = min(with_feat, key=lambda f: OP.fd_min(f["feat"], bank)) if with_feat else pick_best(ffaces, None, qmin, False)
        else:
            bestf = pick_best(ffaces, None, qmin, False)
        faces_local[i] = bestf
        if bestf is not None and bestf.get("quality", 0.0) >= qmin:
            faces_passing_quality += 1
    any_face_visible = faces_passing_quality > 0 if bool(getattr(cfg, "face_visible_uses_quality", True)) else faces_detected > 0
    cands = []
    for i, (x1, y1, x2, y2) in enumerate(boxes):
        bf = faces_local.get(i)
        fd = OP.fd_min(bf["feat"], bank) if (bf is not None and bf.get("feat") is not None and bank is not None) else None
        face_ok = fd is not None and fd <= float(cfg.face_thresh)
        accept = face_ok
        if bool(getattr(cfg, "require_face_if_visible", True)) and any_face_visible and bank is not None:
            qfail = bf is None
            if bf is not None and bf.get("quality", 0.0) < float(getattr(cfg, "face_quality_floor_absurd", 15)):
                qfail = True
            if bf is not None and not face_ok:
                qfail = True
            if qfail:
                accept = False
        if not accept:
            continue
        fb = bf["bbox"]
        fx1 = max(0.0, min(float(W2), float(x1 + fb[0]))); fy1 = max(0.0, min(float(H2), float(y1 + fb[1])))
        fx2 = max(fx1 + 1.0, min(float(W2), float(x1 + fb[2]))); fy2 = max(fy1 + 1.0, min(float(H2), float(y1 + fb[3])))
        cands.append(dict(i=i, box=(x1, y1, x2, y2), fd=fd, score=fd, quality=float(bf.get("quality", 0.0)),
                          face_box=(fx1, fy1, fx2, fy2), area=(x2 - x1) * (y2 - y1), sharp=0.0))
    return cands, dict(faces_detected=faces_detected, faces_passing_quality=faces_passing_quality, any_face_visible=any_face_visible)


def arbitrate(cands, any_face_visible, cfg, lock_hits=0, locked_face=False, prev_box=None, seek_cooldown=0):
    """gui_app.py:7788-7845: ambiguity drop (two face-bearing candidates closer than face_margin_min), order by
    (score, -area, -sharp), keep only the best when the top two scores are within score_margin, then the lock gate
    (fd <= lock_face_thresh and IoU(prev_box, box) >= iou_gate once lock_after_hits hits were saved).  -> chosen or None."""
    cands = list(cands)
    if not cands:
        return None
    if bool(getattr(cfg, "prefer_face_when_available", True)) and any_face_visible:
        fc = sorted([c for c in cands if c.get("fd") is not None], key=lambda d: d["fd"])
        if len(fc) >= 2 and (fc[1]["fd"] - fc[0]["fd"]) < float(getattr(cfg, "face_margin_min", 0.05)):
            return None
    cands.sort(key=lambda c: (c["score"] if c["score"] is not None else 1e9, -c["area"], -c["sharp"]))
    if len(cands) >= 2 and cands[0]["score"] is not None and cands[1]["score"] is not None:
        if abs(cands[0]["score"] - cands[1]["score"]) < float(getattr(cfg, "score_margin", 0.03)):
            cands = cands[:1]
    use_lock = seek_cooldown <= 0 and lock_hits >= int(getattr(cfg, "lock_after_hits", 1)) and locked_face
    for c in cands:
        if use_lock:
            ok = True
            if c.get("fd") is not None:
                ok = ok and c["fd"] <= float(getattr(cfg, "lock_face_thresh", 0.28))
            if prev_box is not None and iou_xyxy(prev_box, c["box"]) < float(getattr(cfg, "iou_gate", 0.05)):
                ok = False
            if not ok:
                continue
        return c
    return cands[0]
