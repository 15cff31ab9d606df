"""Main-pass identity decisions on the GPU path: "is the target in this frame, and where is its face".

Reproduces, for the face-only pipeline (SURVEY.md App. C rules 1-6), the identity kernel of each main-pass call site of
person_capture/gui_app.py: segment gate + seek cooldown (:5648-5684), frame_stride gate (:5743-5746), lock-face ROI probe
(:5796-5855, geometry _expand_xyxy :4186-4199, accept :5919, miss counter :6024-6028), full-frame cadence probe (:6030-6116),
face-only global fallback (:7521-7551), lock-face box update (:7501-7505, :4164-4177), optional runtime bank learning
(:7460-7494), the per-person-crop site for person boxes handed in by the caller (:6269-6346, :6370-6437) and the frame-level
arbitration + lock gate between several candidates (:7788-7845).  The person detector, crop composition and saving are
downstream product logic and out of scope.

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


def box_iou(a, b) -> float:
    """IoU of two xyxy boxes with the reference's 1e-9 guard (gui_app.py:3484-3493)."""
    iw = max(0, min(a[2], b[2]) - max(a[0], b[0]))
    ih = max(0, min(a[3], b[3]) - max(a[1], b[1]))
    inter = iw * ih
    return inter / (max(0, (a[2] - a[0]) * (a[3] - a[1])) + max(0, (b[2] - b[0]) * (b[3] - b[1])) - inter + 1e-9)


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
                    roi = frame[ry1:ry2, rx1:rx2]          # device frames: a view; extract() compacts it on the engine's stream
                    if not isinstance(roi, torch.Tensor):
                        roi = np.ascontiguousarray(roi)
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


def person_site(st: "MainPassIdentity", frame, boxes: Sequence[Tuple[int, int, int, int]]):
    """Per-person-crop identity site (App. C rule 2) for person boxes supplied by the caller (the person detector is out of
    scope).  Every crop -- a device-side view when `frame` is a device tensor -- goes through FaceEmbedder.extract (padded
    retry by face_det_pad when it finds nothing), the crop's face is the one closest to the bank, and a candidate survives
    iff fd <= face_thresh and the "face required if any face is visible" gate lets it through.
    -> (candidates, any_face_visible); candidate = dict(i, box, fd, score, quality, face_box, area, sharp)."""
    cfg, face = st.cfg, st.face
    H2, W2 = int(frame.shape[0]), int(frame.shape[1])
    have_bank = len(st.bank) > 0
    pad = float(getattr(cfg, "face_det_pad", 0.08))
    qmin = float(cfg.face_quality_min)
    thr = float(cfg.face_thresh)

    def crop_of(x1, y1, x2, y2):
        c = frame[y1:y2, x1:x2]                            # device frames: a view; extract() compacts it on the engine's stream
        return c if isinstance(c, torch.Tensor) else np.ascontiguousarray(c)

    def distances(faces):
        if not faces or not have_bank:
            return None
        return st.fd_fn(faces, st.bank.array()) if st.fd_fn is not None else _fds_for_last_faces(face, st.bank)

    best = []            # per box: (face dict in crop coordinates, fd) or None
    n_faces = n_quality = 0
    for (x1, y1, x2, y2) in boxes:
        faces = face.extract(crop_of(x1, y1, x2, y2))
        fds = distances(faces)
        if not faces and pad > 0.0:
            pw, ph = int(round((x2 - x1) * pad)), int(round((y2 - y1) * pad))
            px1, py1, px2, py2 = max(0, x1 - pw), max(0, y1 - ph), min(W2, x2 + pw), min(H2, y2 + ph)
            if (pw > 0 or ph > 0) and px2 > px1 and py2 > py1:
                faces = face.extract(crop_of(px1, py1, px2, py2))
                fds = distances(faces)
                shift = np.array([x1 - px1, y1 - py1, x1 - px1, y1 - py1], np.int32)
                faces = [dict(f, bbox=f["bbox"] - shift) for f in faces]
        n_faces += len(faces)
        pick = None
        if faces:
            k = int(np.argmin(fds)) if fds is not None else faces.index(FaceEmbedder.best_face(faces))
            pick = (faces[k], None if fds is None else float(fds[k]))
            n_quality += int(float(faces[k]["quality"]) >= qmin)
        best.append(pick)
    visible = n_quality > 0 if bool(getattr(cfg, "face_visible_uses_quality", True)) else n_faces > 0
    gate = bool(getattr(cfg, "require_face_if_visible", True)) and visible and have_bank
    floor = float(getattr(cfg, "face_quality_floor_absurd", 15))
    cands = []
    for i, ((x1, y1, x2, y2), pick) in enumerate(zip(boxes, best)):
        if pick is None or pick[1] is None or pick[1] > thr:
            continue                      # no face / no distance / not the target: face_only mode rejects, and so does the gate
        f, fd = pick
        if gate and float(f["quality"]) < floor:
            continue
        fb = f["bbox"]
        fx1 = max(0.0, min(float(W2), float(x1 + fb[0]))); fy1 = max(0.0, min(float(H2), float(y1 + fb[1])))
        fx2 = max(fx1 + 1.0, min(float(W2), float(x1 + fb[2]))); fy2 = max(fy1 + 1.0, min(float(H2), float(y1 + fb[3])))
        cands.append(dict(i=i, box=(x1, y1, x2, y2), fd=fd, score=fd, quality=float(f["quality"]), face_box=(fx1, fy1, fx2, fy2),
                          area=(x2 - x1) * (y2 - y1), sharp=0.0))
    return cands, visible


def arbitrate(cands: Sequence[dict], any_face_visible: bool, cfg, lock_hits: int = 0, locked_face: bool = False, prev_box=None,
              seek_cooldown: int = 0) -> Optional[dict]:
    """Frame-level arbitration between candidates (App. C rules 3-4): drop the frame when the two closest faces are within
    face_margin_min of each other; order by (score, -area, -sharp); a runner-up within score_margin of the winner is
    discarded; once locked (lock_after_hits saved hits, outside the seek cooldown) a candidate must also have
    fd <= lock_face_thresh and overlap the previous box by iou_gate -- if none does, the best one is still taken."""
    if not cands:
        return None
    fds = sorted(c["fd"] for c in cands if c.get("fd") is not None)
    if bool(getattr(cfg, "prefer_face_when_available", True)) and any_face_visible and len(fds) >= 2 \
            and (fds[1] - fds[0]) < float(getattr(cfg, "face_margin_min", 0.05)):
        return None
    order = sorted(cands, key=lambda c: (1e9 if c["score"] is None else c["score"], -c["area"], -c["sharp"]))
    if len(order) >= 2 and None not in (order[0]["score"], order[1]["score"]) \
            and abs(order[0]["score"] - order[1]["score"]) < float(getattr(cfg, "score_margin", 0.03)):
        order = order[:1]
    if seek_cooldown <= 0 and locked_face and lock_hits >= int(getattr(cfg, "lock_after_hits", 1)):
        lthr, gate = float(getattr(cfg, "lock_face_thresh", 0.28)), float(getattr(cfg, "iou_gate", 0.05))
        for c in order:
            if (c.get("fd") is None or c["fd"] <= lthr) and (prev_box is None or box_iou(prev_box, c["box"]) >= gate):
                return c
    return order[0]


def _run_span(st: MainPassIdentity, clip, k: int, spans, fps: float, stride: int, device_frames: bool, log, hits, after_frame=None):
    """Processed frames of kept span k, with the segment-jump rule in front (gui_app.py:5648-5682: a jump to the span start
    arms the seek cooldown).  after_frame(idx) -> True stops the span early (used by the sharded driver's re-runs)."""
    s, e = spans[k]
    prev_next = 0 if k == 0 else spans[k - 1][1] + 1
    if prev_next < s:
        st.cooldown = int(max(2, (fps or 30) * 0.25))
    for idx in range(max(s, prev_next), min(e, clip.total_frames - 1) + 1):
        if idx % stride != 0:
            continue
        frame = clip.device_batch(st.face.engine, [idx])[0] if device_frames else clip.host(idx)
        hit = st.step(idx, frame, log)
        if hit is not None:
            hits.append(hit)
        if after_frame is not None and after_frame(idx):
            return False
    return True


def main_pass(clip, fps: float, keep_spans: Sequence[Tuple[int, int]], face: FaceEmbedder, ref_face_feat, cfg,
              log: Optional[list] = None, device_frames: bool = True, fd_fn=None) -> List[dict]:
    """Sequential main pass over the kept spans of `clip` (a frame source of prescan.py).  -> accepted hits."""
    st = MainPassIdentity(face, ref_face_feat, cfg, fd_fn=fd_fn)
    stride = max(1, int(getattr(cfg, "frame_stride", 2)))
    spans = list(keep_spans) or [(0, clip.total_frames - 1)]
    hits: List[dict] = []
    for k in range(len(spans)):
        _run_span(st, clip, k, spans, fps, stride, device_frames, log, hits)
    return hits


# --------------------------------------------------------------------------------------------------------------------------
# span-sharded main pass (BASELINE config 3 "sharded across 2/4/8 B200")
# --------------------------------------------------------------------------------------------------------------------------
_STATE_LEN = 11


def _get_state(st: MainPassIdentity) -> np.ndarray:
    """Everything the next frame's decisions depend on: lock box / miss counter / seek cooldown of the driver and the
    FaceEmbedder counters that steer the upright size and the adaptive rotation probes."""
    f = st.face
    lb = st.lock_box if st.lock_box is not None else (-1, -1, -1, -1)
    return np.array([st.lock_box is not None, *lb, st.misses, st.cooldown, f._frame_idx, f._no_face_streak,
                     max(f._last_face_idx, -10 ** 9), f._rot_cycle], np.int64)


def _set_state(st: MainPassIdentity, v: np.ndarray):
    f = st.face
    st.lock_box = tuple(int(x) for x in v[1:5]) if v[0] else None
    st.misses, st.cooldown = int(v[5]), int(v[6])
    f._frame_idx, f._no_face_streak, f._last_face_idx, f._rot_cycle = int(v[7]), int(v[8]), int(v[9]), int(v[10])


def _canon(v: np.ndarray, after_hit: int) -> tuple:
    """State up to what can influence behaviour: the absolute extract counter only acts through its distance to the last
    face (saturating at rot_after_hit_frames) and modulo rot_every_n (checked separately through the recorded consults)."""
    return (*[int(x) for x in v[:7]], int(v[8]), int(min(v[7] - v[9], after_hit + 1)))


def _split_blocks(spans, world: int, stride: int):
    """Contiguous blocks of spans with about equal numbers of processed frames: only a block's first span starts from a
    state another rank produced."""
    work = np.array([max(1, (e - s) // stride + 1) for s, e in spans], np.float64)
    cum = np.concatenate([[0.0], np.cumsum(work)])
    cuts = [0]
    for r in range(1, world):
        cuts.append(max(cuts[-1], int(np.searchsorted(cum, cum[-1] * r / world, side="left"))))
    cuts.append(len(spans))
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def main_pass_sharded(clip, fps: float, keep_spans: Sequence[Tuple[int, int]], face: FaceEmbedder, ref_face_feat, cfg,
                      dist_group=None, log: Optional[list] = None, device_frames: bool = True, fd_fn=None, stats: Optional[dict] = None):
    """main_pass with the kept spans split into contiguous blocks, one per rank.  -> the hits of ALL spans on every rank,
    identical to what the sequential main_pass returns.

    A span's decisions depend on the state the previous span left behind (lock-face box, miss counter, FaceEmbedder
    counters -- the reference resets none of them at a segment jump, gui_app.py:5648-5682).  Each rank therefore first
    runs its block from a neutral state, the ranks exchange the states at the block ends (one small all_gather), and a rank
    whose true start state differs re-runs the head of its block from it until its state trajectory meets the first run's
    (a few frames: the next accepted face sets the same lock box) and keeps the rest; a flipped adaptive-rotation decision
    (the only place the absolute extract counter matters) re-runs the block.  Rounds repeat until no end state moves --
    rank r is exact after at most r rounds, in practice after one.  Runtime bank learning makes the bank itself
    sequential state, so learn_bank_runtime falls back to every rank running the sequential pass."""
    import torch.distributed as dist
    world, rank = 1, 0
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(dist_group), dist.get_rank(dist_group)
    spans = list(keep_spans) or [(0, clip.total_frames - 1)]
    if world == 1 or bool(getattr(cfg, "learn_bank_runtime", False)) or len(spans) < 2:
        return main_pass(clip, fps, spans, face, ref_face_feat, cfg, log=log, device_frames=device_frames, fd_fn=fd_fn)
    stride = max(1, int(getattr(cfg, "frame_stride", 2)))
    blocks = _split_blocks(spans, world, stride)
    lo, hi = blocks[rank]
    st = MainPassIdentity(face, ref_face_feat, cfg, fd_fn=fd_fn)
    after_hit, every_n = int(face.rot_after_hit_frames), int(face.rot_every_n)
    neutral = _get_state(st)
    cuda = dist.get_backend(dist_group) == "nccl"
    dev = face.engine.tdev if cuda else torch.device("cpu")

    def run_block(init: np.ndarray, prior=None):
        """-> dict(hits, log, trace {idx: state after the frame}, consults, end).  With `prior` (the previous run of this
        block) the run stops at the first frame whose state equals the prior trajectory and adopts the prior's remainder."""
        _set_state(st, init)
        face.rot_consults = []
        hits, lg, trace = [], [], {}
        merged = None

        def after(idx):
            nonlocal merged
            v = _get_state(st)
            trace[idx] = v
            if prior is not None and idx in prior["trace"]:
                pv = prior["trace"][idx]
                if _canon(v, after_hit) == _canon(pv, after_hit):
                    shift = int(v[7] - pv[7])
                    later = [c + shift for c in prior["consults"] if c > pv[7]]
                    flipped = any(((c + face.rot_phase) % every_n == 0) != ((c - shift + face.rot_phase) % every_n == 0) for c in later)
                    if not flipped:
                        merged = (idx, shift)
                        return True
            return False

        for k in range(lo, hi):
            if not _run_span(st, clip, k, spans, fps, stride, device_frames, lg, hits, after_frame=after):
                break
        consults = list(face.rot_consults)
        end = _get_state(st)
        if merged is not None:
            idx0, shift = merged
            hits += [h for h in prior["hits"] if h["idx"] > idx0]
            lg += [r for r in prior["log"] if r["idx"] > idx0]
            for i, v in prior["trace"].items():
                if i > idx0:
                    w = v.copy()
                    w[7] += shift
                    if w[9] > -10 ** 9:
                        w[9] += shift
                    trace[i] = w
            consults += [c + shift for c in prior["consults"] if c > prior["trace"][idx0][7]]
            end = prior["end"].copy()
            end[7] += shift
            if end[9] > -10 ** 9:
                end[9] += shift
        return dict(hits=hits, log=lg, trace=trace, consults=consults, end=end, init=init.copy())

    def gather_ends(end: np.ndarray) -> np.ndarray:
        mine = torch.as_tensor(end, dtype=torch.int64, device=dev)
        out = torch.empty((world * _STATE_LEN,), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(out, mine, group=dist_group)
        return out.cpu().numpy().reshape(world, _STATE_LEN)

    cur = run_block(neutral) if hi > lo else dict(hits=[], log=[], trace={}, consults=[], end=neutral.copy(), init=neutral.copy())
    rounds = 0
    while True:
        ends = gather_ends(cur["end"])
        # a rank with an empty block hands on what it received
        true_init = neutral.copy()
        for r in range(rank):
            if blocks[r][1] > blocks[r][0]:
                true_init = ends[r]
            # (an empty block's end state equals its init by construction of the next round)
        changed = 0
        if hi > lo and not np.array_equal(true_init, cur["init"]):
            cur = run_block(true_init, prior=cur)
            changed = 1
        elif hi == lo:
            cur["end"], cur["init"] = true_init.copy(), true_init.copy()
        flag = torch.tensor([changed], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.SUM, group=dist_group)
        rounds += 1
        if int(flag.item()) == 0 or rounds > world + 1:
            break
    if stats is not None:
        stats["rounds"], stats["block"], stats["frames_processed"] = rounds, (lo, hi), len(cur["log"])
    # exchange the hits: (idx, site code, fd, quality, face box) per accepted frame, padded to the largest rank
    sites = ("lock_roi", "fullframe", "fallback")
    rows = np.array([[h["idx"], sites.index(h["site"]), h["fd"], h["quality"], *h["face_box"]] for h in cur["hits"]], np.float64).reshape(-1, 8)
    cnt = torch.tensor([len(rows)], dtype=torch.int64, device=dev)
    cnts = torch.empty((world,), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(cnts, cnt, group=dist_group)
    cap = max(int(cnts.max().item()), 1)
    send = torch.zeros((cap, 8), dtype=torch.float64, device=dev)
    if len(rows):
        send[:len(rows)] = torch.as_tensor(rows, dtype=torch.float64, device=dev)
    recv = torch.empty((world * cap * 8,), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send.view(-1), group=dist_group)
    recv = recv.cpu().numpy().reshape(world, cap, 8)
    out = []
    for r in range(world):
        for row in recv[r, :int(cnts[r].item())]:
            out.append(dict(idx=int(row[0]), site=sites[int(row[1])], fd=float(row[2]), quality=float(row[3]),
                            face_box=tuple(int(v) for v in row[4:8])))
    if log is not None:
        log.extend(cur["log"])          # this rank's frames only
    return out


INDEX_COLUMNS = ["frame", "time_secs", "score", "face_dist", "reid_dist", "x1", "y1", "x2", "y2", "crop_path", "sharpness", "ratio"]


def index_rows(hits: Sequence[dict], fps: float) -> List[list]:
    """index.csv rows (gui_app.py:5148) as far as the identity path determines them: frame, time, score = face distance
    (face-only pipeline: no ReID distance), and the accepted face box.  crop_path / sharpness / ratio belong to the crop
    composition and saving stages (out of scope) and stay empty."""
    return [[h["idx"], h["idx"] / float(fps or 30.0), h["fd"], h["fd"], "", *h["face_box"], "", "", ""] for h in hits]


def fullframe_identity(clip, idxs: Sequence[int], face: FaceEmbedder, ref_face_feat, cfg, batch: int = 16, max_faces: int = 4096):
    """Throughput form of the full-frame identity site (normal mode: flip-TTA on) for the frames `idxs`, with the results of
    calling FaceEmbedder.extract(frame, imgsz=face_fullframe_imgsz) on them one after the other.

    Phase 1 (batched): upright SCRFD pass at round32(face_fullframe_imgsz), K4 align, ArcFace e(x)+e(flip x) in 444-image
    runs, bank distances.  Phase 2 (host, in frame order): the FaceEmbedder counters are advanced exactly as extract() would
    advance them, and the frames the batched pass cannot answer are handed to extract() itself -- a frame whose upright
    pass found nothing (the reference then walks its scale-TTA / pad-probe / adaptive-rotation chain, face_embedder.py:
    2251-2433) and a frame that follows three empty ones (the upright size drops to fast_no_face_imgsz, :2193-2194).
    -> list of dict(idx, n_faces, fd, accept, face_box, quality, via)."""
    from .face_embedder import round32, _MIN_SIDE
    from .prescan import FaceTable
    eng = face.engine
    bank = RefBank(cfg, ref_face_feat)
    eng.set_bank(bank.array())
    imgsz = getattr(cfg, "face_fullframe_imgsz", None)
    imgsz = int(imgsz) if imgsz else None
    dyn = round32(max(_MIN_SIDE, imgsz or 640))
    qmin = float(cfg.face_quality_min)
    use_qv = bool(getattr(cfg, "face_visible_uses_quality", True))
    thr = float(cfg.face_thresh)
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
    fds_all = np.zeros((0,), np.float64)
    if table.count:
        _, sim, _ = eng.match(table.flip, None, None, table.count, want_feat=False)
        eng.sync()
        fds_all = 1.0 - sim[:table.count].cpu().numpy().astype(np.float64)

    def decide(rec, boxes, qual, fds, W2, H2):
        sel = sorted(range(len(boxes)), key=lambda i: (qual[i], (boxes[i][2] - boxes[i][0]) * (boxes[i][3] - boxes[i][1])), reverse=True)
        cand = [i for i in sel if qual[i] >= qmin] if use_qv else sel      # the reference sorts by (quality, area) first: ties go to the first
        g = min(cand or sel, key=lambda i: fds[i])
        fx1, fy1, fx2, fy2 = [float(v) for v in boxes[g]]
        fx1 = max(0.0, min(float(W2), fx1)); fy1 = max(0.0, min(float(H2), fy1))
        fx2 = max(fx1 + 1.0, min(float(W2), fx2)); fy2 = max(fy1 + 1.0, min(float(H2), fy2))
        x1 = max(0, min(W2 - 1, int(round(fx1)))); y1 = max(0, min(H2 - 1, int(round(fy1))))
        rec.update(fd=float(fds[g]), accept=bool(fds[g] <= thr), quality=float(qual[g]),
                   face_box=(x1, y1, max(x1 + 1, min(W2, int(round(fx2)))), max(y1 + 1, min(H2, int(round(fy2))))))

    out = []
    for chunk, H2, W2, counts, rows, boxes, qual in metas:
        off = 0
        for b, idx in enumerate(chunk):
            k = int(counts[b])
            rec = dict(idx=idx, n_faces=k, fd=None, accept=False, face_box=None, quality=None, via="batched")
            if face._no_face_streak >= 3 and min(dyn, face.fast_no_face_imgsz) != dyn or k == 0:
                # the sequential call sees a different upright size, or has to walk the empty-frame chain: let it
                faces = face.extract(clip.device_batch(eng, [idx])[0], imgsz=imgsz)
                rec.update(n_faces=len(faces), via="extract")
                if faces and len(bank):
                    fds = _fds_for_last_faces(face, bank)
                    decide(rec, [f["bbox"] for f in faces], [f["quality"] for f in faces], fds, W2, H2)
            else:
                face._frame_idx += 1                      # what extract() does when its upright pass finds faces
                face._no_face_streak, face._last_face_idx, face._rot_cycle = 0, face._frame_idx, 0
                decide(rec, boxes[off:off + k], qual[off:off + k], fds_all[rows[off:off + k]], W2, H2)
            out.append(rec)
            off += k
    return out
