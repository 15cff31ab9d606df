"""Main-pass identity decisions on the GPU path: "is the target in this frame, and where is its face".

Reproduces, for the face-only pipeline, the identity kernel of each main-pass call site of
person_capture/gui_app.py: segment gate + seek cooldown (:5648-5684), frame_stride gate (:5743-5746), lock-face ROI probe
(:5796-5855, geometry _expand_xyxy :4186-4199, accept :5919, miss counter :6024-6028), full-frame cadence probe (:6030-6116),
face-only global fallback (:7521-7551), lock-face box update (:7501-7505, :4164-4177), optional runtime bank learning
(:7460-7494).  Person association, crop composition and saving are downstream product logic and out of scope.

Two drivers:
  * main_pass            -- the reference's sequential loop (the lock-face ROI of frame i depends on the accept of frame i-1);
                            every extract / distance runs in libpcb200, ROI crops are device-side views of resident frames.
  * fullframe_identity   -- throughput form of the full-frame site for a batch of frames (BASELINE config 3): one batched
                            SCRFD pass at `face_fullframe_imgsz`, flip-TTA embeddings, bank distances and the accept rule.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .face_embedder import FaceEmbedder
from .prescan import RefBank, _fds_for_last_faces


def expand_xyxy(box, pad_x: float, pad_y: float, frame_w: int, frame_h: int) -> Tuple[int, int, int, int]:
    x1, y1, x2, y2 = [float(v) for v in box]
    ix1 = max(0, min(frame_w - 1, int(math.floor(x1 - pad_x))))
    iy1 = max(0, min(frame_h - 1, int(math.floor(y1 - pad_y))))
    return ix1, iy1, max(ix1 + 1, min(frame_w, int(math.ceil(x2 + pad_x)))), max(iy1 + 1, min(frame_h, int(math.ceil(y2 + pad_y))))


def _argmin_face(faces: Sequence[dict], fds: np.ndarray, quality_min: float, use_quality_vis: bool) -> Optional[int]:
    """Index of the face the reference calls `gbest`: min fd over quality-passing faces, else over all (first wins ties)."""
    if not len(faces):
        return None
    idx = list(range(len(faces)))
    if use_quality_vis:
        passing = [i for i in idx if float(faces[i]["quality"]) >= quality_min]
        idx = passing or idx
    return min(idx, key=lambda i: fds[i])


class MainPassIdentity:
    """State of the identity side of Processor.run's per-frame loop."""

    def __init__(self, face: FaceEmbedder, ref_face_feat, cfg, fd_fn=None):
        """`fd_fn(faces, bank_array) -> fd per returned face` replaces the GPU matcher in CPU tests of the host logic."""
        self.face, self.cfg, self.fd_fn = face, cfg, fd_fn
        self.bank = RefBank(cfg, ref_face_feat)
        self.lock_box: Optional[Tuple[int, int, int, int]] = None
        self.misses = 0
        self.cooldown = 0
        self.last_add = -10 ** 9

    # one identity site: extract on `img` (host array or device tensor) + gbest + accept
    def _site(self, img, name: str, off: Tuple[int, int], W2: int, H2: int, rec: dict):
        cfg = self.cfg
        imgsz = getattr(cfg, "face_fullframe_imgsz", None)
        faces = self.face.extract(img, imgsz=int(imgsz) if imgsz is not None else None)
        rec["n_faces"] = max(rec["n_faces"], len(faces))
        if not faces or not len(self.bank):
            return None
        fds = self.fd_fn(faces, self.bank.array()) if self.fd_fn is not None else _fds_for_last_faces(self.face, self.bank)
        g = _argmin_face(faces, fds, float(cfg.face_quality_min), bool(getattr(cfg, "face_visible_uses_quality", True)))
        fd = float(fds[g])
        rec["fd"] = fd
        if fd > float(cfg.face_thresh):
            return None
        fx1, fy1, fx2, fy2 = [float(v) for v in faces[g]["bbox"]]
        fx1 += off[0]; fx2 += off[0]; fy1 += off[1]; fy2 += off[1]
        fx1 = max(0.0, min(float(W2), fx1)); fy1 = max(0.0, min(float(H2), fy1))
        fx2 = max(fx1 + 1.0, min(float(W2), fx2)); fy2 = max(fy1 + 1.0, min(float(H2), fy2))
        return dict(site=name, fd=fd, feat=faces[g]["feat"], quality=float(faces[g]["quality"]), box=(fx1, fy1, fx2, fy2))

    def step(self, idx: int, frame, log: Optional[list] = None) -> Optional[dict]:
        """One processed frame (host ndarray or device uint8 tensor [H,W,3]).  -> hit dict or None."""
        cfg = self.cfg
        H2, W2 = int(frame.shape[0]), int(frame.shape[1])
        rec = dict(idx=idx, site=None, roi=None, n_faces=0, fd=None, accept=False)
        cand = None
        have_bank = len(self.bank) > 0
        if have_bank and bool(getattr(cfg, "lock_face_roi_enable", True)) and self.lock_box is not None and self.cooldown <= 0:
            ran = False
            lx1, ly1, lx2, ly2 = [float(v) for v in self.lock_box]
            if lx2 > 0.0 and ly2 > 0.0 and lx1 < float(W2) and ly1 < float(H2):
                pad = max(0.0, float(getattr(cfg, "lock_face_roi_pad", 1.25)))
                rx1, ry1, rx2, ry2 = expand_xyxy((lx1, ly1, lx2, ly2), max(16.0, max(1.0, lx2 - lx1) * pad),
                                                 max(16.0, max(1.0, ly2 - ly1) * pad), W2, H2)
                if rx2 > rx1 + 8 and ry2 > ry1 + 8:
                    ran = True
                    rec["roi"] = (rx1, ry1, rx2, ry2)
                    roi = frame[ry1:ry2, rx1:rx2]
                    roi = roi.contiguous() if isinstance(roi, torch.Tensor) else np.ascontiguousarray(roi)
                    cand = self._site(roi, "lock_roi", (rx1, ry1), W2, H2, rec)
                    if cand is not None:
                        self.misses = 0
            if ran and cand is None:
                self.misses += 1
                if self.misses > max(0, int(getattr(cfg, "lock_face_roi_max_misses", 8))):
                    self.lock_box, self.misses = None, 0
        when_missed = bool(getattr(cfg, "face_fullframe_when_missed", True))
        if cand is None and have_bank and when_missed:
            cadence = int(getattr(cfg, "face_fullframe_cadence", 12))
            if cadence <= 0 or (idx % max(1, cadence) == 0):
                cand = self._site(frame, "fullframe", (0, 0), W2, H2, rec)
        if cand is None and have_bank and when_missed:
            cand = self._site(frame, "fallback", (0, 0), W2, H2, rec)
        hit = None
        if cand is not None:
            x1, y1, x2, y2 = cand["box"]
            fx1 = max(0, min(W2 - 1, int(round(x1)))); fy1 = max(0, min(H2 - 1, int(round(y1))))
            fx2 = max(fx1 + 1, min(W2, int(round(x2)))); fy2 = max(fy1 + 1, min(H2, int(round(y2))))
            rec.update(site=cand["site"], fd=cand["fd"], accept=True)
            hit = dict(idx=idx, site=cand["site"], fd=cand["fd"], quality=cand["quality"], face_box=(fx1, fy1, fx2, fy2))
            if bool(getattr(cfg, "learn_bank_runtime", False)):
                qmin = float(cfg.face_quality_min)
                cd = int(getattr(cfg, "prescan_add_cooldown_samples", 5)) * max(1, int(getattr(cfg, "frame_stride", 2)))
                if cand["fd"] <= float(getattr(cfg, "prescan_fd_add", 0.22)) and (qmin <= 0 or cand["quality"] >= qmin) \
                        and (idx - self.last_add) >= cd:
                    if self.bank.offer(cand["feat"], cand["quality"]) in ("added", "replaced"):
                        self.last_add = idx
            if fx2 > fx1 and fy2 > fy1:
                self.lock_box, self.misses = (fx1, fy1, fx2, fy2), 0
        if log is not None:
            log.append(rec)
        if self.cooldown > 0:
            self.cooldown -= 1
        return hit


def main_pass(clip, fps: float, keep_spans: Sequence[Tuple[int, int]], face: FaceEmbedder, ref_face_feat, cfg,
              log: Optional[list] = None, device_frames: bool = True, fd_fn=None) -> List[dict]:
    """Sequential main pass over the kept spans of `clip` (a frame source of prescan.py).  -> accepted hits."""
    st = MainPassIdentity(face, ref_face_feat, cfg, fd_fn=fd_fn)
    stride = max(1, int(getattr(cfg, "frame_stride", 2)))
    total = clip.total_frames
    spans = list(keep_spans)
    hits: List[dict] = []
    span_i, frame_idx = 0, 0
    jump_cd = int(max(2, (fps or 30) * 0.25))
    while frame_idx < total:
        if spans:
            if span_i >= len(spans):
                break
            s, e = spans[span_i]
            if frame_idx < s:
                frame_idx, st.cooldown = s, jump_cd
                continue
            if frame_idx > e:
                span_i += 1
                if span_i >= len(spans):
                    break
                if frame_idx < spans[span_i][0]:
                    frame_idx, st.cooldown = spans[span_i][0], jump_cd
                continue
        idx = frame_idx
        frame_idx += 1
        if idx % stride != 0:
            continue
        frame = clip.device_batch(face.engine, [idx])[0] if device_frames else clip.host(idx)
        hit = st.step(idx, frame, log)
        if hit is not None:
            hits.append(hit)
    return hits


def fullframe_identity(clip, idxs: Sequence[int], face: FaceEmbedder, ref_face_feat, cfg, batch: int = 16, max_faces: int = 4096):
    """Throughput form of the full-frame identity site (normal mode: flip-TTA on) for the frames `idxs`.

    One batched upright SCRFD pass at round32(face_fullframe_imgsz), K4 align, ArcFace e(x)+e(flip x), bank distance, gbest
    + accept rule per frame.  Frames whose upright pass finds no face are returned with n_faces = 0 (the reference would walk
    its scale-TTA / pad-probe / rotation chain for them, which is stateful; callers that need it run FaceEmbedder.extract on
    those frames).  -> list of dict(idx, n_faces, fd, accept, face_box, quality)."""
    from .face_embedder import round32, _MIN_SIDE
    eng = face.engine
    bank = RefBank(cfg, ref_face_feat)
    eng.set_bank(bank.array())
    imgsz = getattr(cfg, "face_fullframe_imgsz", None)
    dyn = round32(max(_MIN_SIDE, int(imgsz) if imgsz else 640))
    qmin = float(cfg.face_quality_min)
    use_qv = bool(getattr(cfg, "face_visible_uses_quality", True))
    thr = float(cfg.face_thresh)
    # phase 1: SCRFD + K4 per frame batch; chips are queued and embedded in 444-image runs (the ArcFace batch must not
    # depend on how few faces a frame batch holds)
    from .prescan import FaceTable
    table = FaceTable(lazy=False)
    metas = []
    for b0 in range(0, len(idxs), batch):
        chunk = list(idxs[b0:b0 + batch])
        frames = clip.device_batch(eng, chunk)
        n, H2, W2, _ = frames.shape
        det = eng.detect(frames, dyn, face.conf, min_box=int(face.scrfd_min_box_px), max_det=face.max_det)
        al = eng.align(frames, det, max_faces=max_faces)
        eng.sync()
        total = int(al.face_total.cpu()[0])
        counts = al.face_count.cpu().numpy()
        rows = table.queue(eng, al.chips, total) if total else np.zeros((0,), np.int64)
        metas.append((chunk, H2, W2, counts, rows, al.face_box[:total].cpu().numpy() if total else None,
                      al.quality[:total].cpu().numpy() if total else None))
    table.finalize(eng)
    # phase 2: distances of normalise(e(x) + e(flip x)) against the bank, then gbest + accept per frame
    fds_all = np.zeros((0,), np.float64)
    if table.count:
        _, sim, _ = eng.match(table.flip, None, None, table.count, want_feat=False)
        eng.sync()
        fds_all = 1.0 - sim[:table.count].cpu().numpy().astype(np.float64)
    out = []
    for chunk, H2, W2, counts, rows, boxes, qual in metas:
        fds = fds_all[rows] if len(rows) else None
        off = 0
        for b, idx in enumerate(chunk):
            k = int(counts[b])
            rec = dict(idx=idx, n_faces=k, fd=None, accept=False, face_box=None, quality=None)
            if k:
                sel = list(range(off, off + k))
                # the reference sorts faces by (quality, area) before taking the argmin, which decides ties
                sel.sort(key=lambda i: (qual[i], (boxes[i][2] - boxes[i][0]) * (boxes[i][3] - boxes[i][1])), reverse=True)
                cand = [i for i in sel if qual[i] >= qmin] if use_qv else sel
                g = min(cand or sel, key=lambda i: fds[i])
                fx1, fy1, fx2, fy2 = [float(v) for v in boxes[g]]
                fx1 = max(0.0, min(float(W2), fx1)); fy1 = max(0.0, min(float(H2), fy1))
                fx2 = max(fx1 + 1.0, min(float(W2), fx2)); fy2 = max(fy1 + 1.0, min(float(H2), fy2))
                x1 = max(0, min(W2 - 1, int(round(fx1)))); y1 = max(0, min(H2 - 1, int(round(fy1))))
                rec.update(fd=float(fds[g]), accept=bool(fds[g] <= thr), quality=float(qual[g]),
                           face_box=(x1, y1, max(x1 + 1, min(W2, int(round(fx2)))), max(y1 + 1, min(H2, int(round(fy2))))))
            out.append(rec)
            off += k
    return out
