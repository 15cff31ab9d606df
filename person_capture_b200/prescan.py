"""Pre-scan driver: kept-span detection over a clip with the identity path on the GPU.

Reference behaviour reproduced (person_capture/gui_app.py):
  _fd_min :660-674 (on the GPU: pcb_match), _stream_ref_bank_update :922-986, reference-bank
  build :4517-4556, _prescan :1140-1222 + :1468-1655 (stride sampling, fd9 gate, hysteresis,
  pad/clamp/min-len/merge), bridge :1657-1668, _refine_edges :1671-1830, cache :787-920.
Out of scope: UI command queue, decode/seek, Qt status, preview (frames come from a FrameSource).

Two drivers produce the same spans/bank:
  * prescan_sequential -- the reference's loop verbatim in structure: one FaceEmbedder.extract per
    sample (GPU kernels, host policy), distances from pcb_match.
  * prescan_batched    -- the throughput path: the GPU computes, for batches of
    samples, a state-independent *superset* (upright pass; for upright-empty samples both rotated
    probes and the heavy passes they trigger; e(x) and e(flip x) for every face), then the host
    replays the reference's sequential state machine over those records, with distances against
    the live bank recomputed on the GPU whenever the bank changes.  With torch.distributed, ranks
    take contiguous time chunks, all-gather the per-face records (NCCL) and replay identically.
The wall-clock refine budget of the reference (gui_app.py:1692-1696) is not reproduced
(non-deterministic): refinement always completes.
"""
from __future__ import annotations

import ast
import hashlib
import json
import itertools
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .face_embedder import FaceEmbedder, _ROT_PAD

FD_NONE = 9.0


# --------------------------------------------------------------------------------------------
# frame sources
# --------------------------------------------------------------------------------------------
class HostClip:
    """Frames produced on the host (decode stand-in): get(i) -> uint8 BGR [H,W,3]."""
    host_resident = True     # batches cross PCIe: compute_superset prefetches the next one on the engine's copy stream

    def __init__(self, get: Callable[[int], np.ndarray], total_frames: int):
        self._get = get
        self.total_frames = int(total_frames)

    def host(self, i: int) -> np.ndarray:
        return self._get(i)

    def device_batch(self, eng, idxs: Sequence[int], stream=None) -> torch.Tensor:
        return eng.to_device(np.stack([self._get(i) for i in idxs]), stream=stream)


class VideoFileClip:
    """Frames decoded on the HOST by OpenCV's FFmpeg reader -- the reference's default SDR reader
    (`cv2.VideoCapture(video_path, cv2.CAP_FFMPEG)`, gui_app.py:4667-4721; its pre-scan walks it with grab / retrieve,
    :1417-1464).  Each batch is decoded straight into a pinned buffer and copied to HBM on the copy stream, so the
    host-resident pre-scan (prefetch three batches ahead, ArcFace on the second context) applies unchanged.  Decode itself stays
    on the CPU (a few ms per 1080p frame): this is the functional bridge from a file to the identity path, not a fast one -- an
    NVDEC source is not built.  Sequential access costs one `read` per frame; any other index seeks first."""
    host_resident = True

    def __init__(self, path: str):
        import cv2
        self._cv2 = cv2
        self.path = str(path)
        self.cap = cv2.VideoCapture(self.path, cv2.CAP_FFMPEG)
        if not self.cap.isOpened():
            self.cap = cv2.VideoCapture(self.path)
        if not self.cap.isOpened():
            raise L.PcbError(f"cannot open video {self.path!r}")
        self.total_frames = int(self.cap.get(cv2.CAP_PROP_FRAME_COUNT) or 0)
        self.fps = float(self.cap.get(cv2.CAP_PROP_FPS) or 0.0)
        self.width = int(self.cap.get(cv2.CAP_PROP_FRAME_WIDTH) or 0)
        self.height = int(self.cap.get(cv2.CAP_PROP_FRAME_HEIGHT) or 0)
        if self.total_frames <= 0 or self.width <= 0 or self.height <= 0:
            raise L.PcbError(f"video {self.path!r} reports no frames / no size")
        self._next = 0
        self.decoded = 0                # frames decoded so far (sequential reads and seeks alike)
        self.seeks = 0

    def _read_into(self, i: int, out: Optional[np.ndarray]) -> np.ndarray:
        if i < 0 or i >= self.total_frames:
            raise IndexError(i)
        if i != self._next:
            self.cap.set(self._cv2.CAP_PROP_POS_FRAMES, int(i))
            self.seeks += 1
        ok, frame = self.cap.read() if out is None else self.cap.read(out)
        if not ok or frame is None:
            raise L.PcbError(f"video {self.path!r}: frame {i} of {self.total_frames} could not be decoded")
        self._next = i + 1
        self.decoded += 1
        if out is not None and frame is not out:
            np.copyto(out, frame)
            return out
        return frame

    def host(self, i: int) -> np.ndarray:
        return self._read_into(int(i), None)

    def device_batch(self, eng, idxs: Sequence[int], stream=None) -> torch.Tensor:
        # a fresh pinned tensor per batch: torch's host allocator recycles the block only after the copy that reads it has
        # completed on its stream, so batches prefetched ahead never overwrite one another
        buf = torch.empty((len(idxs), self.height, self.width, 3), dtype=torch.uint8, pin_memory=True)
        view = buf.numpy()
        for k, i in enumerate(idxs):
            self._read_into(int(i), view[k])
        with torch.cuda.stream(stream if stream is not None else eng.stream):
            return buf.to(eng.tdev, non_blocking=True)

    def close(self):
        if self.cap is not None:
            self.cap.release()
            self.cap = None


class DeviceClip:
    """Frames already resident in HBM as a uint8 tensor [N,H,W,3] (north_star: pre-decoded in HBM)."""

    def __init__(self, frames: torch.Tensor, first_index: int = 0):
        assert frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4
        self.frames = frames
        self.first = int(first_index)
        self.total_frames = self.first + frames.shape[0]

    def host(self, i: int) -> np.ndarray:
        return self.frames[i - self.first].cpu().numpy()

    def device_batch(self, eng, idxs: Sequence[int], stream=None) -> torch.Tensor:
        lo = idxs[0] - self.first
        if list(idxs) == list(range(idxs[0], idxs[0] + len(idxs))):
            return self.frames[lo:lo + len(idxs)]
        with torch.cuda.stream(eng.stream):
            sel = torch.as_tensor([i - self.first for i in idxs], device=self.frames.device)
            return self.frames.index_select(0, sel)


# --------------------------------------------------------------------------------------------
# bank (host side: a few 512-d dot products per accepted face, as in the reference)
# --------------------------------------------------------------------------------------------
def _weights3(cfg) -> Tuple[float, float, float]:
    w = getattr(cfg, "prescan_weights", (0.70, 0.25, 0.05))
    if isinstance(w, str):
        txt = w.strip()
        if txt:
            try:
                w = json.loads(txt)
            except Exception:
                try:
                    w = ast.literal_eval(txt)
                except Exception:
                    w = (0.70, 0.25, 0.05)
    try:
        if isinstance(w, (list, tuple)) and len(w) >= 3:
            return float(w[0]), float(w[1]), float(w[2])
    except Exception:
        pass
    return 0.70, 0.25, 0.05


_BANK_SERIAL = itertools.count(1)


class RefBank:
    """The live reference bank (rows are unit vectors) with the reference's streaming update."""

    def __init__(self, cfg, rows: Optional[np.ndarray] = None):
        self.cfg = cfg
        # identity for Engine.set_bank's "already uploaded" check.  NOT id(self): CPython hands the address of a freed bank to
        # the next one, and a new bank (version 0) then looked like the previous main pass's bank -- distances were computed
        # against a stale device bank (found by test_main_pass_on_device_frames_is_repeatable_* after another main-pass test)
        self.serial = next(_BANK_SERIAL)
        self.rows: List[np.ndarray] = []
        if rows is not None:
            arr = np.asarray(rows, dtype=np.float32)
            if arr.ndim == 1:
                arr = arr.reshape(1, -1)
            arr = arr / np.maximum(np.linalg.norm(arr, axis=1, keepdims=True), 1e-6)
            self.rows = [r.copy() for r in arr]
        self.version = 0

    def array(self) -> Optional[np.ndarray]:
        if not self.rows:
            return None
        if getattr(self, "_arr_version", None) != self.version or getattr(self, "_arr", None) is None or len(self._arr) != len(self.rows):
            self._arr = np.ascontiguousarray(np.vstack(self.rows), np.float32)
            self._arr_version = self.version
        return self._arr

    def __len__(self):
        return len(self.rows)

    def offer(self, vec, quality: float) -> str:
        """-> 'skip' | 'added' | 'dup' | 'replaced' (gui_app.py:922-986)."""
        if vec is None:
            return "skip"
        cfg = self.cfg
        cap = max(1, int(getattr(cfg, "prescan_bank_max", 64)))
        dedup = float(getattr(cfg, "prescan_diversity_dedup_cos", 0.968))
        margin = float(getattr(cfg, "prescan_replace_margin", 0.010))
        wa, wd, wq = _weights3(cfg)
        v = np.asarray(vec, dtype=np.float32).reshape(-1)
        nv = float(np.linalg.norm(v))
        if nv <= 1e-6:
            return "skip"
        v = v / nv
        if not self.rows:
            self.rows.append(v)
            self.version += 1
            return "added"
        B = self.array()              # cached float32 stack of the rows (same values np.vstack(...).astype(float32) gives)
        sims = B @ v
        top = float(sims.max())
        if top >= dedup:
            return "dup"
        if len(self.rows) < cap:
            self.rows.append(v)
            self.version += 1
            return "added"
        cos_a = max(-1.0, min(1.0, float(np.dot(B[0], v))))
        s_new = wa * (1.0 - float(np.sqrt(max(0.0, 2.0 - 2.0 * cos_a)))) + wd * (1.0 - top) \
            + wq * float(min(max(quality or 0.0, 0.0), 1000.0) / 300.0)
        G = B @ B.T
        np.fill_diagonal(G, -1.0)
        ca = np.clip(B @ B[0], -1.0, 1.0)
        s_bank = wa * (1.0 - np.sqrt(np.maximum(0.0, 2.0 - 2.0 * ca))) + wd * (1.0 - G.max(axis=1))
        worst = int(np.argmin(s_bank))
        if s_new > float(s_bank[worst]) + margin:
            self.rows[worst] = v
            self.version += 1
            return "replaced"
        return "skip"


def build_reference_bank(face: FaceEmbedder, ref_images: Sequence[np.ndarray], cfg) -> Optional[np.ndarray]:
    """gui_app.py:4517-4556: every reference image and its mirror, best face, streaming update."""
    bank = RefBank(cfg)
    for img in ref_images:
        for aug in (img, np.ascontiguousarray(img[:, ::-1])):
            bf = FaceEmbedder.best_face(face.extract(aug))
            if bf is not None and bf.get("feat") is not None:
                bank.offer(bf["feat"], float(bf.get("quality", 0.0)))
    return bank.array()


# --------------------------------------------------------------------------------------------
# span state machine (gui_app.py:1468-1655)
# --------------------------------------------------------------------------------------------
class SpanTracker:
    def __init__(self, cfg, fps: int, total_frames: int):
        self.cfg = cfg
        self.fps = fps
        self.total = int(total_frames)
        self.stride = max(1, int(cfg.prescan_stride))
        self.pad = int(round(cfg.prescan_pad_sec * fps))
        self.min_len = int(round(cfg.prescan_min_segment_sec * fps))
        self.enter = float(cfg.prescan_fd_enter)
        self.exit = float(cfg.prescan_fd_exit)
        self.exit_cool = int(round(max(0.0, float(getattr(cfg, "prescan_exit_cooldown_sec", 0.5))) * fps))
        self.spans: List[Tuple[int, int]] = []
        self.active = False
        self.start = 0
        self.neg_run = 0
        self.fd9_streak = 0
        self._fd9_skip = bool(getattr(cfg, "prescan_fd9_skip", True))
        self._fd9_grace = max(0, int(getattr(cfg, "prescan_fd9_grace", 1)))
        self._fd9_period = max(1, int(getattr(cfg, "prescan_fd9_probe_period", 2)))

    def gate_skips(self) -> bool:
        """fd9 skip gate evaluated before a sample (gui_app.py:1479-1492)."""
        if self.active or not self._fd9_skip:
            return False
        return self.fd9_streak >= self._fd9_grace and (self.fd9_streak % self._fd9_period) != 0

    def _close(self, s: int, e: int):
        if e - s + 1 >= self.min_len:
            if self.spans and s <= self.spans[-1][1] + 1:
                self.spans[-1] = (self.spans[-1][0], max(self.spans[-1][1], e))
            else:
                self.spans.append((s, e))

    def observe(self, idx: int, best: float):
        self.fd9_streak = self.fd9_streak + 1 if best >= 8.99 else 0
        if best <= self.enter:
            if not self.active:
                self.active = True
                self.fd9_streak = 0
                self.start = idx
            self.neg_run = 0
        elif self.active:
            self.neg_run += 1
            if self.neg_run * self.stride >= self.exit_cool or best >= self.exit:
                self._close(max(0, self.start - self.pad), min(self.total - 1, idx + self.pad))
                self.active = False
                self.neg_run = 0
                self.fd9_streak = 0

    def finish(self) -> List[Tuple[int, int]]:
        if self.active:
            self._close(max(0, self.start - self.pad), self.total - 1)
        return list(self.spans)


def bridge_spans(spans, gap: int):
    if not spans:
        return spans
    out = []
    cs, ce = spans[0]
    for s, e in spans[1:]:
        if s - ce <= gap:
            ce = max(ce, e)
        else:
            out.append((cs, ce))
            cs, ce = s, e
    out.append((cs, ce))
    return out


def sample_indices(total_frames: int, stride: int) -> List[int]:
    return list(range(0, total_frames, max(1, stride)))


# --------------------------------------------------------------------------------------------
# face-side runtime configuration (gui_app.py:1162-1194, 1858-1868)
# --------------------------------------------------------------------------------------------
def _clamped_conf(cfg) -> float:
    try:
        return min(0.95, max(0.01, float(getattr(cfg, "prescan_face_conf", 0.5))))
    except Exception:
        return 0.5


class _PrescanFaceMode:
    def __init__(self, face: FaceEmbedder, cfg):
        self.face, self.cfg = face, cfg

    def __enter__(self):
        f, cfg = self.face, self.cfg
        self.saved = (f.conf, f.rot_adaptive)
        f.conf = _clamped_conf(cfg)
        f._probe_conf = float(getattr(cfg, "prescan_probe_conf", 0.03))
        f._prescan_period = int(getattr(cfg, "prescan_rot_probe_period", 3))
        f._prescan_probe_imgsz = int(getattr(cfg, "prescan_probe_imgsz", 512))
        f._prescan_no_upscale_det = bool(getattr(cfg, "prescan_no_upscale_det", True))
        f._high_90 = int(getattr(cfg, "prescan_heavy_90", 1536))
        f._high_180 = int(getattr(cfg, "prescan_heavy_180", 1280))
        f.configure_rotation_strategy(adaptive=False)
        f.set_prescan_fast(True, mode="rr")
        f.set_prescan_hint(escalate=False)
        return f

    def __exit__(self, *exc):
        f = self.face
        f.configure_rotation_strategy(adaptive=bool(self.saved[1]))
        f.set_prescan_fast(False)
        f.set_prescan_hint(escalate=False)
        f.conf = self.saved[0]
        return False


def _downscaled_dims(h: int, w: int, wmax: int) -> Tuple[int, int]:
    nh = int(round(h * (wmax / float(w))))
    return nh, wmax


def _maybe_downscale_host(face: FaceEmbedder, frame: np.ndarray, wmax: int, guard_positive: bool):
    """The INTER_AREA pre-scan downscale (gui_app.py:1505-1507; refine adds `Wmax > 0`, :1744) on the GPU."""
    h, w = frame.shape[:2]
    if w > wmax and (wmax > 0 or not guard_positive):
        eng = face.engine
        if guard_positive:
            sc = float(wmax) / float(w)
            nh, nw = int(round(h * sc)), int(round(w * sc))
        else:
            nh, nw = _downscaled_dims(h, w, wmax)
        out = eng.resize(eng.to_device(frame[None]), nh, nw, area=True)
        eng.sync()
        return out[0].cpu().numpy()
    return frame


# --------------------------------------------------------------------------------------------
# sequential driver
# --------------------------------------------------------------------------------------------
def _fds_for_last_faces(face: FaceEmbedder, bank: RefBank) -> np.ndarray:
    """fd of every face returned by the last extract() against `bank`, computed by pcb_match."""
    eng = face.engine
    feats = face.last_feats_dev
    f = face.last_face_count
    eng.set_bank(bank.array(), token=(bank.serial, bank.version))
    _, sim, _ = eng.match(feats, None, None, f, want_feat=False)
    eng.sync()
    return 1.0 - sim[:f].cpu().numpy().astype(np.float64)


def prescan_sequential(clip, fps: int, face: FaceEmbedder, ref_feat, cfg, log: Optional[list] = None):
    """-> (spans, bank).  `fps` is int(round(fps)) as the reference passes it (gui_app.py:5049)."""
    total = clip.total_frames
    bank = RefBank(cfg, ref_feat)
    trk = SpanTracker(cfg, fps, total)
    wmax = int(getattr(cfg, "prescan_max_width", 0))
    fd_add = float(getattr(cfg, "prescan_fd_add", trk.enter))
    cooldown = int(getattr(cfg, "prescan_add_cooldown_samples", 5))
    last_add = -10 ** 9
    with _PrescanFaceMode(face, cfg):
        for sample_idx, idx in enumerate(sample_indices(total, trk.stride)):
            face._prescan_rr_mode = "full" if trk.active else "rr"
            face.set_prescan_hint(escalate=trk.active)
            best = FD_NONE
            skipped = trk.gate_skips()
            nfaces = 0
            if not skipped:
                frame = _maybe_downscale_host(face, clip.host(idx), wmax, guard_positive=False)
                faces = face.extract(frame)
                nfaces = len(faces)
                if faces:
                    fds = _fds_for_last_faces(face, bank)
                    for j, f in enumerate(faces):
                        fd = float(fds[j])
                        best = min(best, fd)
                        if fd <= fd_add and (sample_idx - last_add) >= cooldown and f["quality"] >= cfg.face_quality_min:
                            if bank.offer(f["feat"], float(f["quality"])) in ("added", "replaced"):
                                last_add = sample_idx
                                fds = _fds_for_last_faces(face, bank)   # later faces see the updated bank
            if log is not None:
                log.append(dict(idx=idx, skip=skipped, best=best, active_before=trk.active, nfaces=nfaces))
            trk.observe(idx, best)
        spans = trk.finish()
        spans = _post_process(spans, clip, fps, face, bank, ref_feat, cfg, trk, wmax)
    out_bank = bank.array()
    return spans, (out_bank if out_bank is not None else ref_feat)


def _post_process(spans, clip, fps, face, bank: RefBank, ref_feat, cfg, trk: SpanTracker, wmax: int, batched: int = 0, stats=None,
                  shard=None, known=None):
    gap = int(round(cfg.prescan_bridge_gap_sec * fps))
    do_bridge = getattr(cfg, "prescan_bridge_gap_sec", 0) > 0
    if spans and do_bridge:
        spans = bridge_spans(spans, gap)
    if batched:
        spans = _refine_edges_batched(spans, clip, fps, face, bank, ref_feat, cfg, trk, batched, stats=stats, shard=shard, known=known)
    else:
        spans = _refine_edges(spans, clip, fps, face, bank, ref_feat, cfg, trk, wmax)
    if spans and do_bridge:
        spans = bridge_spans(spans, gap)
    return spans


def _refine_edges(spans, clip, fps, face, bank: RefBank, ref_feat, cfg, trk: SpanTracker, wmax: int):
    """gui_app.py:1671-1830 without the wall-clock budget."""
    if not spans:
        return spans
    total = clip.total_frames
    stride_ref = max(1, min(int(max(1, cfg.prescan_stride) // 4), int(getattr(cfg, "prescan_refine_stride_min", 3))))
    win = int(round(max(0.0, float(getattr(cfg, "prescan_boundary_refine_sec", 0.75))) * fps))
    search = max(int(round(max(0.0, float(cfg.prescan_pad_sec)) * fps)), win)
    trim = bool(getattr(cfg, "prescan_trim_pad", True))
    skip_trailing = bool(getattr(cfg, "prescan_skip_trailing_refine", True))
    rr_old = face._prescan_rr_mode
    face._prescan_rr_mode = "full"
    face.set_prescan_hint(escalate=True)
    use_bank = bank if len(bank) else RefBank(cfg, ref_feat)

    def hit(j: int) -> bool:
        frame = _maybe_downscale_host(face, clip.host(j), wmax, guard_positive=True)
        faces = face.extract(frame)
        if not faces:
            return False
        return bool((_fds_for_last_faces(face, use_bank) <= trk.enter).any())

    out = []
    for s, e in spans:
        ls, le = s, e
        first = None
        j = s
        while j <= min(e, s + search):
            if hit(j):
                first = j
                break
            j += stride_ref
        if first is not None and trim:
            ls = max(s, first)
        last = None
        if not (skip_trailing and e >= total - 1):
            j = max(ls, e - search)
            while j <= e:
                if hit(j):
                    last = j
                j += stride_ref
        if last is not None and trim:
            le = min(e, last)
        if le >= ls and (le - ls + 1) >= trk.min_len:
            out.append((ls, le))
    face.set_prescan_hint(escalate=False)
    face._prescan_rr_mode = rr_old
    return out


def _refine_edges_batched(spans, clip, fps, face, bank: RefBank, ref_feat, cfg, trk: SpanTracker, batch: int, stats=None, shard=None,
                          known=None):
    """Same result as _refine_edges, but every candidate probe frame of all spans goes through the
    batched superset (two GPU rounds: all left windows, then all right windows, whose start depends
    on the refined left edge).  Probes run in "full"/escalate mode: flip-TTA on, 90 then 270.

    `known` = dict(pos={frame: sample index}, meta=flat records, table=face table of the main scan): a probe frame that was a
    SAMPLE of the main scan is not detected / embedded again -- its record is the same function of the same frame (the
    superset is state independent), so the probe reduces to the flip-TTA distance of its faces to the final bank (missing flip
    features are computed for those rows only).  With stride 1 every probe frame is such a frame."""
    if not spans:
        return spans
    total = clip.total_frames
    stride_ref = max(1, min(int(max(1, cfg.prescan_stride) // 4), int(getattr(cfg, "prescan_refine_stride_min", 3))))
    win = int(round(max(0.0, float(getattr(cfg, "prescan_boundary_refine_sec", 0.75))) * fps))
    search = max(int(round(max(0.0, float(cfg.prescan_pad_sec)) * fps)), win)
    trim = bool(getattr(cfg, "prescan_trim_pad", True))
    skip_trailing = bool(getattr(cfg, "prescan_skip_trailing_refine", True))
    use_bank = bank if len(bank) else RefBank(cfg, ref_feat)

    def evaluate(frame_ids):
        """-> {frame: bool hit} for probe frames in escalate mode against the final bank."""
        all_ids = sorted(set(frame_ids))
        if not all_ids:
            return {}
        res_known = {}
        if known is not None:
            pos, meta, t = known["pos"], known["meta"], known["table"]
            rows_of = {}
            for j in all_ids:
                if j not in pos:
                    continue
                m = meta[pos[j]]
                c0 = 0 if m[0] >= 0 else next((6 + 2 * d for d in (0, 1) if m[2 + d] and m[4 + d] and m[6 + 2 * d] >= 0), None)
                rows_of[j] = np.arange(m[c0], m[c0] + m[c0 + 1]) if c0 is not None else None
            need = [r for r in rows_of.values() if r is not None]
            fdf = None
            if need:
                if getattr(t, "lazy", False) and t.ensure_flip(face.engine, np.concatenate(need)):
                    known["dist"].invalidate()          # (several ranks: every rank asks for the same rows -- a matched collective)
                _, fdf = known["dist"].get(use_bank)
            res_known = {j: bool(r is not None and (fdf[r] <= trk.enter).any()) for j, r in rows_of.items()}
            all_ids = [j for j in all_ids if j not in res_known]
            if not all_ids:
                return res_known
        # several ranks: each probe frame is evaluated by the rank that owns its time chunk, results are all-gathered
        ids = all_ids if shard is None else [j for j in all_ids if shard["owner"](j) == shard["rank"]]
        res = {}
        if ids:
            records, table = compute_superset(clip, ids, face, cfg, batch=batch)
            _count_passes(stats, table)
            fdp, fdf = _LiveDistances(face.engine, table).get(use_bank)
        for j in ids:
            rec = records[j]
            chosen = rec.up
            if chosen is None:
                for deg in (90, 270):
                    if rec.hits.get(deg, 0) and rec.heavy_raw.get(deg, 0) and deg in rec.heavy:
                        chosen = rec.heavy[deg]
                        break
            res[j] = bool(chosen is not None and (fdf[chosen.rows] <= trk.enter).any())
        if shard is not None:
            # every rank knows which probe frames every other rank owns: one all_gather of padded hit bytes
            import torch.distributed as dist
            world = shard["world"]
            owned = [[j for j in all_ids if shard["owner"](j) == r] for r in range(world)]
            cap = max(max(len(o) for o in owned), 1)
            cuda = dist.get_backend(shard["group"]) == "nccl"
            dev = face.engine.tdev if cuda else torch.device("cpu")
            mine = torch.zeros((cap,), dtype=torch.uint8)
            for k, j in enumerate(owned[shard["rank"]]):
                mine[k] = 1 if res[j] else 0
            recv = torch.empty((world, cap), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(recv.view(-1), mine.to(dev), group=shard["group"])
            hits = recv.cpu().numpy()
            res = {j: bool(hits[r, k]) for r in range(world) for k, j in enumerate(owned[r])}
        res.update(res_known)
        return res

    left_ids = []
    for s, e in spans:
        left_ids += list(range(s, min(e, s + search) + 1, stride_ref))
    # the right window starts at max(refined left edge, e - search); the refined left edge is at most s + search, so for
    # spans with e - search >= s + search the window does not depend on the left result and both go in ONE GPU round
    early_right = []
    for s, e in spans:
        if not (skip_trailing and e >= total - 1) and e - search >= s + search:
            early_right += list(range(e - search, e + 1, stride_ref))
    hit = evaluate(left_ids + early_right)
    lefts = []
    for s, e in spans:
        first = next((j for j in range(s, min(e, s + search) + 1, stride_ref) if hit[j]), None)
        lefts.append(max(s, first) if (first is not None and trim) else s)
    right_ids = []
    for (s, e), ls in zip(spans, lefts):
        if not (skip_trailing and e >= total - 1):
            right_ids += [j for j in range(max(ls, e - search), e + 1, stride_ref) if j not in hit]
    rhit = dict(hit)
    rhit.update(evaluate(right_ids))
    out = []
    for (s, e), ls in zip(spans, lefts):
        le = e
        if not (skip_trailing and e >= total - 1):
            hits = [j for j in range(max(ls, e - search), e + 1, stride_ref) if rhit[j]]
            if hits and trim:
                le = min(e, hits[-1])
        if le >= ls and (le - ls + 1) >= trk.min_len:
            out.append((ls, le))
    return out


# --------------------------------------------------------------------------------------------
# batched driver: GPU superset + host replay
# --------------------------------------------------------------------------------------------
@dataclass
class _Variant:
    """Faces of one sample from one candidate final pass, in kept order."""
    box: np.ndarray        # [k,4] int32
    quality: np.ndarray    # [k] float64
    rows: np.ndarray       # [k] row indices into the face table


@dataclass
class SampleRecord:
    idx: int
    up: Optional[_Variant] = None
    hits: Dict[int, int] = field(default_factory=dict)        # rotation -> probe hits
    heavy_raw: Dict[int, int] = field(default_factory=dict)   # rotation -> heavy-pass NMS count
    heavy: Dict[int, _Variant] = field(default_factory=dict)


class _Timeline:
    """Optional GPU timeline of a pre-scan step (PCB_TIMELINE=file): named CUDA events per stream, dumped as milliseconds
    from the first one.  Diagnostics only."""
    path = os.environ.get("PCB_TIMELINE")

    def __init__(self):
        self.marks = []

    def mark(self, name: str, stream):
        if self.path:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            self.marks.append((name, ev))

    def dump(self):
        if self.path and self.marks:
            torch.cuda.synchronize()
            t0 = self.marks[0][1]
            with open(self.path, "a") as fh:
                fh.write(json.dumps([(n, round(t0.elapsed_time(e), 3)) for n, e in self.marks]) + "\n")
        self.marks = []


TIMELINE = _Timeline()


class EarlyFlips:
    """Flip-TTA passes issued WHILE the superset is still running.  When frames arrive from the host the SMs wait for PCIe
    between frame batches (timeline, profiles/r02_e2e_timeline.txt: a 64-frame batch is 3.5 ms of SCRFD per 7-9 ms of
    copy), so a host-resident pre-scan embeds its chips in SMALL runs as they accumulate and follows each plain run with the
    flip pass the replay will most likely need, all on a SECOND context (own stream): an ArcFace run queued on the main
    stream would sit behind the next SCRFD pass, which itself waits for its frames -- on its own stream it runs in that
    wait.  (First attempts, measured: flips alone on a second lower-priority context -- priorities do not preempt
    persistent kernels and the first 444-chip run only existed after the copies had ended: no gain; small runs + flips on
    the main stream: 116 -> 112 ms, the SMs still idle 16 of the first 32 ms.)  Which rows: the rule of `_predict_flip_rows` (a sample
    that follows, within the exit cooldown, a sample with a face at plain distance <= enter + margin from the initial
    bank), applied run by run with the distances known so far.  A prediction only: the exact prediction after the superset
    and the on-demand path of the replay compute whatever is missing, so results never depend on it."""

    def __init__(self, engine, cfg, fps: int, carry_in: bool, margin: float = 0.12):
        self.engine = engine        # the context the ArcFace runs of the superset go to (its own stream: see FaceTable.flush)
        self.thr = float(cfg.prescan_fd_enter) + margin
        stride = max(1, int(cfg.prescan_stride))
        exit_cool = int(round(max(0.0, float(getattr(cfg, "prescan_exit_cooldown_sec", 0.5))) * fps))
        self.tail = (exit_cool + stride - 1) // stride + 1
        self.hot = np.zeros((0,), np.int64)          # sorted sample positions with a face within the threshold
        if carry_in:                                  # the chunk may start inside another rank's active stretch
            self.hot = np.array([-1], np.int64)

    def select(self, sample_pos: np.ndarray, fd0: np.ndarray) -> np.ndarray:
        """-> indices (into this run) of the rows to flip."""
        hot_now = np.unique(sample_pos[fd0 <= self.thr])
        if len(hot_now):
            self.hot = np.union1d(self.hot, hot_now)
        if not len(self.hot):
            return np.zeros((0,), np.int64)
        k = np.searchsorted(self.hot, sample_pos, side="left")          # hot samples strictly before the row's sample
        prev = np.where(k > 0, self.hot[np.maximum(k - 1, 0)], -10 ** 9)
        return np.nonzero(sample_pos - prev <= self.tail)[0].astype(np.int64)


class FaceTable:
    """All faces of the superset: normalised features without / with flip-TTA (device + host).

    Chips are queued as the frame batches are aligned and embedded in runs of EMBED_RUN images, so the ArcFace batch
    (which sets the wave efficiency of the convolution tiles) does not depend on how many faces a frame batch holds.
    Lazy mode (single process): only e(x) is computed up front, as the reference does while no span is active
    (face_embedder.py:1295); chips and raw embeddings stay resident and `ensure_flip` computes e(flip x) for the rows the
    replay actually evaluates in the active state (plus a look-ahead window, so the GPU sees large batches)."""
    # images per ArcFace graph run (= pcb_embed's chunk; half as many faces when both variants are computed): whole waves of
    # tiles in the 14x14 stage -- 504 images with the trailing-pad layout of the small maps (443 pair tiles = 5.99 waves on 74
    # CTA pairs; 28x28: 11.97 waves, 56x56: 44.8), 444 with the ring everywhere (PCB_SMALL_PAD_MAX=0)
    EMBED_RUN = int(os.environ.get("PCB_EMBED_RUN", "504" if int(os.environ.get("PCB_SMALL_PAD_MAX", "16")) >= 14 else "444"))
    # host-resident clips (early flips on): plain runs of this many images as the chips accumulate, and flip passes of at least
    # EARLY_RUN images (148 images = one full wave of 14x14 pair tiles)
    HOST_RUN = int(os.environ.get("PCB_HOST_RUN", "252"))
    FIRST_RUN = int(os.environ.get("PCB_FIRST_RUN", "74"))     # the first run goes out as soon as this many chips exist
    EARLY_RUN = int(os.environ.get("PCB_EARLY_RUN", "148"))

    def __init__(self, lazy: bool = False, early: Optional["EarlyFlips"] = None):
        self.lazy = lazy
        self.early = early if lazy else None   # flip passes issued while the superset runs (host-resident clips)
        if self.early is not None:
            self.EMBED_RUN = min(self.EMBED_RUN, self.HOST_RUN)
        self.sp_parts: List[np.ndarray] = []   # per row: position of its sample in the rank's sample list (early flips only)
        self.runs: List[tuple] = []            # lazy mode, per flushed run: (first row, rows, sim to the device bank, event)
        self.runs_decided = 0
        self.early_rows: List[np.ndarray] = [] # rows selected for an early flip pass, not issued yet
        self.early_n = 0
        self.early_done: List[tuple] = []      # (rows, normalised flip features on the device, completion event)
        self.feat_plain: List[torch.Tensor] = []
        self.feat_flip: List[torch.Tensor] = []
        self.raw: List[torch.Tensor] = []
        self.chip_list: List[torch.Tensor] = []
        self.pending: List[torch.Tensor] = []
        self.pending_n = 0
        self.count = 0
        self.q_parts: List[np.ndarray] = []    # per-row quality / box area in row order (flat records for pcb_replay)
        self.a_parts: List[np.ndarray] = []
        self.encoded = None
        self.flip_passes = 0          # faces that went through the flip pass (bench: ArcFace image passes / s)

    def queue(self, eng, chips: torch.Tensor, k: int, sample_pos: Optional[np.ndarray] = None) -> np.ndarray:
        """Register k aligned chips; -> their row numbers.  Embedding happens in `flush`."""
        rows = np.arange(self.count, self.count + k)
        if self.early is not None:
            self.sp_parts.append(np.asarray(sample_pos if sample_pos is not None else np.full(k, -10 ** 9), np.int64))
        # early mode: the caller has waited for K4 (the chips are final), and the copy goes to the ArcFace context's stream --
        # on the main stream it would sit behind the next SCRFD pass, which is already queued and waits for its frames
        st = self._work_stream(eng)
        if st is not eng.stream:
            chips.record_stream(st)
        with torch.cuda.stream(st):
            self.pending.append(chips[:k].clone())
        self.pending_n += k
        self.count += k
        per_run = self.EMBED_RUN if self.lazy else self.EMBED_RUN // 2
        if self.early is not None and not self.runs and self.pending_n >= self.FIRST_RUN:
            self.flush(eng)                 # host-resident start-up: the SMs have nothing else to do yet, embed what is there
        elif self.pending_n >= per_run:
            self.flush(eng, keep_remainder=True)
        return rows

    def _work_stream(self, eng):
        return self.early.engine.stream if self.early is not None else eng.stream

    def flush(self, eng, keep_remainder: bool = False):
        if not self.pending_n:
            return
        per_run = self.EMBED_RUN if self.lazy else self.EMBED_RUN // 2
        with torch.cuda.stream(self._work_stream(eng)):
            chips = torch.cat(self.pending, 0) if len(self.pending) > 1 else self.pending[0]
        n = chips.shape[0]
        take = (n // per_run) * per_run if keep_remainder else n
        if take == 0:
            return
        with torch.cuda.stream(self._work_stream(eng)):
            use = chips[:take].contiguous()
        if self.lazy:
            aeng = self.early.engine if self.early is not None else eng
            TIMELINE.mark(f"arc{len(self.runs)}+", aeng.stream)
            emb, _ = aeng.embed(use, take, False)
            fp, sim, _ = aeng.match(emb, None, None, take)        # normalise(e(x)); sim: against the bank on that context
            TIMELINE.mark(f"arc{len(self.runs)}-", aeng.stream)
            self.raw.append(emb[:take])
            self.chip_list.append(use)
            if self.early is not None:
                ev = torch.cuda.Event()
                ev.record(aeng.stream)
                self.runs.append((self.count - self.pending_n, take, sim, ev))
                self._early_step(eng)
        else:
            emb, emb_flip = eng.embed(use, take, True)
            fp, _, _ = eng.match(emb, None, None, take)           # normalise(e(x))
            ff, _, _ = eng.match(emb, emb_flip, None, take)       # normalise(e(x) + e(flip x))
            self.feat_flip.append(ff[:take])
            self.flip_passes += take
        self.feat_plain.append(fp[:take])
        if take < n:
            with torch.cuda.stream(self._work_stream(eng)):
                self.pending = [chips[take:].clone()]
            self.pending_n = n - take
        else:
            self.pending, self.pending_n = [], 0

    # ---- early flips: see EarlyFlips
    def _early_step(self, eng):
        """Decide, for every flushed run whose similarities have arrived, which of its rows will most likely need the flip
        feature, and issue flip passes for them on the second context.  Never waits for the GPU."""
        ef = self.early
        sp_all = None
        while self.runs_decided < len(self.runs):
            row0, take, sim, ev = self.runs[self.runs_decided]
            if not ev.query():
                break
            if sp_all is None:
                sp_all = np.concatenate(self.sp_parts)
            fd0 = 1.0 - sim[:take].cpu().numpy().astype(np.float64)
            rows = ef.select(sp_all[row0:row0 + take], fd0) + row0
            if len(rows):
                self.early_rows.append(rows)
                self.early_n += len(rows)
            self.runs_decided += 1
        while self.early_n >= self.EARLY_RUN:
            self._issue_early(eng, min(self.early_n, 444))

    def _issue_early(self, eng, n: int):
        feng = self.early.engine        # the stream the plain runs went to: the pass follows the runs it belongs to
        allr = np.concatenate(self.early_rows)
        rows, rest = allr[:n], allr[n:]
        self.early_rows = [rest] if len(rest) else []
        self.early_n = len(rest)
        starts = np.asarray([r[0] for r in self.runs], np.int64)
        part = np.searchsorted(starts, rows, side="right") - 1
        order = []
        with torch.cuda.stream(feng.stream):
            chips, raws = [], []
            for pi in np.unique(part):
                feng.stream.wait_event(self.runs[pi][3])           # the run's chips / raw embeddings are final
                sel = rows[part == pi]
                loc = torch.as_tensor(sel - starts[pi], device=self.chip_list[pi].device)
                chips.append(self.chip_list[pi].index_select(0, loc))
                raws.append(self.raw[pi].index_select(0, loc))
                order.append(sel)
            chips = torch.cat(chips, 0).contiguous()
            raws = torch.cat(raws, 0).contiguous()
        order = np.concatenate(order)
        TIMELINE.mark(f"flip{len(self.early_done)}({len(order)})+", feng.stream)
        _, emb_flip = feng.embed(chips, len(order), "only")
        ff, _, _ = feng.match(raws, emb_flip, None, len(order))
        TIMELINE.mark(f"flip{len(self.early_done)}-", feng.stream)
        done = torch.cuda.Event()
        done.record(feng.stream)
        self.early_done.append((order, ff[:len(order)], done, chips, raws))      # inputs stay referenced until the pass has run
        self.flip_passes += len(order)

    def finalize(self, eng):
        self.flush(eng)
        if self.early is not None and self.runs:
            self.runs[-1][3].synchronize()      # the superset ends here anyway: decide the last runs and issue what is left
            self._early_step(eng)
            if self.early_n:
                self._issue_early(eng, self.early_n)
            self.early.engine.sync()            # everything the second context produced is final before the main stream reads it
        with torch.cuda.stream(eng.stream):
            if self.count:
                self.plain = torch.cat(self.feat_plain, 0).contiguous()
                if self.lazy:
                    self.raw_all = torch.cat(self.raw, 0).contiguous()
                    self.chips = torch.cat(self.chip_list, 0).contiguous()
                    self.flip = torch.zeros_like(self.plain)
                else:
                    self.flip = torch.cat(self.feat_flip, 0).contiguous()
            else:
                self.plain = eng.empty((1, L.FEAT_DIM), torch.float32)
                self.flip = eng.empty((1, L.FEAT_DIM), torch.float32)
        self.flip_ready = np.zeros(self.count, bool) if self.lazy else np.ones(self.count, bool)
        self.flip_host = np.zeros((self.count, L.FEAT_DIM), np.float32) if self.lazy else None
        if self.early_done:
            # flips computed on the second context while the superset was running
            for rows, ff, done, _c, _r in self.early_done:
                eng.stream.wait_event(done)
                with torch.cuda.stream(eng.stream):
                    self.flip.index_copy_(0, torch.as_tensor(rows, device=self.flip.device), ff)
            eng.sync()
            for rows, ff, _d, _c, _r in self.early_done:
                self.flip_host[rows] = ff.cpu().numpy()
                self.flip_ready[rows] = True
            self.early_done = []
        self.feat_plain, self.feat_flip, self.raw, self.chip_list = [], [], [], []

    def ensure_flip(self, eng, rows: np.ndarray) -> bool:
        """Compute normalise(e(x) + e(flip x)) for the rows that do not have it yet.  -> True if anything was computed."""
        if not self.lazy or not len(rows):
            return False
        need = np.unique(np.asarray(rows)[~self.flip_ready[rows]])
        if not len(need):
            return False
        with torch.cuda.stream(eng.stream):
            sel = torch.as_tensor(need, device=self.plain.device)
            chips = self.chips.index_select(0, sel).contiguous()
            raw = self.raw_all.index_select(0, sel).contiguous()
        _, emb_flip = eng.embed(chips, len(need), "only")
        ff, _, _ = eng.match(raw, emb_flip, None, len(need))
        with torch.cuda.stream(eng.stream):
            self.flip.index_copy_(0, sel, ff[:len(need)])
        eng.sync()
        self.flip_host[need] = ff[:len(need)].cpu().numpy()
        self.flip_ready[need] = True
        self.flip_passes += len(need)
        return True


def _collect_variant(eng, frames, det, idx_list, records, key, table: FaceTable, max_faces: int, al=None, pos_of=None):
    """Align every accumulated face of `det` (batch over idx_list) and queue its chip for embedding.  `al`: the align
    result if the caller already enqueued it and waited for it."""
    if al is None:
        al = eng.align(frames, det, max_faces=max_faces)
        eng.sync()
    total = int(al.face_total.cpu()[0])
    counts = al.face_count.cpu().numpy()
    if total == 0:
        return
    sample_pos = None
    if table.early is not None and pos_of is not None:
        sample_pos = np.repeat(np.asarray([pos_of[i] for i in idx_list], np.int64), counts[:len(idx_list)])
    rows = table.queue(eng, al.chips, total, sample_pos)
    boxes = al.face_box[:total].cpu().numpy()
    qual = al.quality[:total].cpu().numpy()
    b64 = boxes.astype(np.int64)
    table.q_parts.append(qual.astype(np.float64))
    table.a_parts.append((b64[:, 2] - b64[:, 0]) * (b64[:, 3] - b64[:, 1]))
    off = 0
    for b, sidx in enumerate(idx_list):
        k = int(counts[b])
        if k:
            v = _Variant(boxes[off:off + k].astype(np.int32), qual[off:off + k].astype(np.float64), rows[off:off + k])
            if key == "up":
                records[sidx].up = v
            else:
                records[sidx].heavy[key] = v
        off += k


def compute_superset(clip, idxs: Sequence[int], face: FaceEmbedder, cfg, batch: int = 32, max_faces: int = 4096,
                     lazy_flip: bool = False, early: Optional[EarlyFlips] = None):
    """GPU stage of the batched pre-scan for the samples `idxs` (fast pre-scan settings must be active).
    lazy_flip: compute e(flip x) later, only for the faces the replay evaluates while a span is active.

    Software-pipelined over frame batches: the upright SCRFD pass and K4 of batch k+1 are enqueued before the host waits
    (on an event, not a stream sync) for batch k's counts, so the SMs never wait for the host's bookkeeping; host frames of
    batch k+2 cross PCIe meanwhile on the copy stream."""
    eng = face.engine
    wmax = int(getattr(cfg, "prescan_max_width", 0))
    records: Dict[int, SampleRecord] = {}
    table = FaceTable(lazy=lazy_flip, early=early)
    pos_of = {idx: k for k, idx in enumerate(idxs)} if early is not None else None
    chunks = [list(idxs[b0:b0 + batch]) for b0 in range(0, len(idxs), batch)]
    prefetch = bool(getattr(clip, "host_resident", False))

    def fetch(ci):
        """Issue the H2D copy of batch ci on the copy stream; -> (frames, event that marks its completion)."""
        if not prefetch or ci >= len(chunks):
            return None
        with torch.cuda.stream(eng.copy_stream):
            TIMELINE.mark(f"copy{ci}+", eng.copy_stream)
            fr = clip.device_batch(eng, chunks[ci], stream=eng.copy_stream)
            ev = torch.cuda.Event()
            ev.record(eng.copy_stream)
            TIMELINE.mark(f"copy{ci}-", eng.copy_stream)
        return fr, ev

    # host frames: the copy stream runs PREFETCH batches ahead of the SMs.  The link, not the SMs, paces a host-resident
    # pre-scan, so it must never idle: with one batch of lead, every burst of extra work on the SMs (an early flip pass, the
    # rotated passes of an empty batch) stalled the copies behind it and that time was lost for good.
    depth = max(1, int(os.environ.get("PCB_PREFETCH", "3")))
    fetched = {ci: fetch(ci) for ci in range(min(depth, len(chunks)))} if prefetch else {}

    def issue(ci):
        """Enqueue K0 + upright SCRFD + K4 of batch ci; nothing here waits for the GPU."""
        chunk = chunks[ci]
        if prefetch:
            frames, ev = fetched.pop(ci)
            eng.stream.wait_event(ev)
            frames.record_stream(eng.stream)
            if ci + depth < len(chunks):
                fetched[ci + depth] = fetch(ci + depth)
        else:
            frames = clip.device_batch(eng, chunk)
        n, h, w, _ = frames.shape
        if w > wmax:
            nh, nw = _downscaled_dims(h, w, wmax)
            frames = eng.resize(frames, nh, nw, area=True)
            h, w = nh, nw
        dyn = face.upright_size(h, w, None)
        TIMELINE.mark(f"det{ci}+", eng.stream)
        det0 = eng.detect(frames, dyn, face.conf, min_box=int(face.scrfd_min_box_px), max_det=face.max_det)
        al = eng.align(frames, det0, max_faces=max_faces)
        TIMELINE.mark(f"det{ci}-", eng.stream)
        done = torch.cuda.Event()
        done.record(eng.stream)
        return dict(chunk=chunk, frames=frames, h=h, w=w, dyn=dyn, det0=det0, al=al, done=done)

    meta_parts: List[np.ndarray] = []

    def encode_chunk(chunk):
        # flat pcb_replay records of a finished chunk; runs while the next batch is on the SMs
        m = np.zeros((len(chunk), L.REPLAY_META), np.int32)
        m[:, 0] = m[:, 6] = m[:, 8] = -1
        for s_i, idx in enumerate(chunk):
            rec = records[idx]
            if rec.up is not None:
                m[s_i, 0], m[s_i, 1] = rec.up.rows[0], len(rec.up.rows)
            if rec.hits or rec.heavy:
                for d, deg in enumerate((90, 270)):
                    m[s_i, 2 + d] = rec.hits.get(deg, 0)
                    m[s_i, 4 + d] = rec.heavy_raw.get(deg, 0)
                    if deg in rec.heavy:
                        m[s_i, 6 + 2 * d], m[s_i, 7 + 2 * d] = rec.heavy[deg].rows[0], len(rec.heavy[deg].rows)
        meta_parts.append(m)

    cur = issue(0) if chunks else None
    for ci in range(len(chunks)):
        nxt = issue(ci + 1) if ci + 1 < len(chunks) else None
        cur["done"].synchronize()
        chunk, frames, dyn, det0 = cur["chunk"], cur["frames"], cur["dyn"], cur["det0"]
        n = frames.shape[0]
        heavy90, heavy180 = face.heavy_sizes(cur["h"], cur["w"], dyn)
        for i in chunk:
            records[i] = SampleRecord(i)
        acc0 = det0.acc_count.cpu().numpy()
        _collect_variant(eng, frames, det0, chunk, records, "up", table, max_faces, al=cur["al"], pos_of=pos_of)
        if table.early is not None:
            table._early_step(eng)          # flips of the runs whose distances have arrived go out now
        empty = [b for b in range(n) if acc0[b] == 0]
        cur = nxt
        if not empty:
            encode_chunk(chunk)
            continue
        with torch.cuda.stream(eng.stream):
            sub = frames.index_select(0, torch.as_tensor(empty, device=frames.device)).contiguous()
        sub_idx = [chunk[b] for b in empty]
        probes = {deg: eng.detect(sub, face.probe_size(dyn), face.probe_conf_value(), rot=deg, max_det=face.max_det)
                  for deg in (90, 270)}
        eng.sync()
        for deg in (90, 270):
            hits = probes[deg].raw_count.cpu().numpy()
            for j, sidx in enumerate(sub_idx):
                records[sidx].hits[deg] = int(hits[j])
            sel = [j for j in range(len(sub_idx)) if hits[j] > 0]
            if not sel:
                continue
            with torch.cuda.stream(eng.stream):
                sub2 = sub.index_select(0, torch.as_tensor(sel, device=sub.device)).contiguous()
            sel_idx = [sub_idx[j] for j in sel]
            hv = eng.detect(sub2, face.heavy_size_fast(deg, heavy90, heavy180), face.rotated_conf(deg), rot=deg, pad=_ROT_PAD,
                            fix_mode=L.FIX_UNPAD, min_box=0, max_det=face.max_det)
            eng.sync()
            raw = hv.raw_count.cpu().numpy()
            for j, sidx in enumerate(sel_idx):
                records[sidx].heavy_raw[deg] = int(raw[j])
            _collect_variant(eng, sub2, hv, sel_idx, records, deg, table, max_faces, pos_of=pos_of)
        encode_chunk(chunk)
    table.finalize(eng)
    one = lambda parts, dt: np.concatenate(parts).astype(dt) if parts else np.zeros(1, dt)
    table.encoded = (np.concatenate(meta_parts, 0) if meta_parts else np.zeros((0, L.REPLAY_META), np.int32),
                     one(table.q_parts, np.float64), one(table.a_parts, np.int64))
    return records, table


class _LiveDistances:
    """fd of every table row against the live bank, for both flip variants; recomputed on the GPU
    (pcb_match) whenever the bank version changes or new flip rows became available.  Both variants go through
    ONE match launch over the stacked [plain; flip] table and one device->host copy."""

    def __init__(self, eng, table: FaceTable):
        self.eng, self.table = eng, table
        self.version = None
        self.fd_plain = self.fd_flip = None
        self.both = None
        self.sim_dev = self.arg_dev = self.sim_host = None
        self.refreshes = 0

    def invalidate(self):
        self.version = None
        self.both = None

    def get(self, bank: RefBank):
        if self.version != bank.version or self.fd_plain is None:
            eng, t = self.eng, self.table
            if t.count == 0:
                self.fd_plain = self.fd_flip = np.zeros((0,), np.float64)
            else:
                if self.both is None:
                    with torch.cuda.stream(eng.stream):
                        # rows without a flip feature yet are zero vectors: their distances are never read
                        self.both = torch.cat([t.plain[:t.count], t.flip[:t.count]], 0).contiguous()
                self.refreshes += 1
                n2 = 2 * t.count
                if self.sim_dev is None or self.sim_dev.shape[0] < n2:
                    self.sim_dev = eng.empty((n2,), torch.float32)
                    self.arg_dev = eng.empty((n2,), torch.int32)
                    self.sim_host = _pinned("live_sim", (n2,), torch.float32)
                eng.set_bank(bank.array())
                eng._check(eng.lib.pcb_match(eng.ctx, self.both.data_ptr(), None, None, n2, None, self.sim_dev.data_ptr(),
                                             self.arg_dev.data_ptr()), "pcb_match")
                with torch.cuda.stream(eng.stream):
                    self.sim_host[:n2].copy_(self.sim_dev[:n2], non_blocking=True)
                eng.sync()      # one synchronisation per bank change
                fd = 1.0 - self.sim_host[:n2].numpy().astype(np.float64)
                self.fd_plain, self.fd_flip = fd[:t.count], fd[t.count:]
            self.version = bank.version
        return self.fd_plain, self.fd_flip


def _count_passes(stats, table: FaceTable):
    if stats is not None:
        stats["faces"] = stats.get("faces", 0) + table.count
        stats["arcface_passes"] = stats.get("arcface_passes", 0) + table.count + table.flip_passes


def _rows_of(rec: "SampleRecord") -> List[np.ndarray]:
    out = [] if rec.up is None else [rec.up.rows]
    return out + [v.rows for v in rec.heavy.values()]


def _replay_python(records: Dict[int, SampleRecord], table: FaceTable, feats_host, idxs: Sequence[int], fps: int, total_frames: int,
                   face: FaceEmbedder, ref_feat, cfg, log: Optional[list] = None, distances=None):
    """Host replay of the reference's sequential loop over precomputed superset records.
    `distances` (tests only) replaces the GPU matcher with an object exposing get(bank).

    Every rank of a multi-GPU pre-scan runs this over ALL samples, so its cost per sample bounds the scaling: the common
    case (no face of the sample can be offered to the bank) is a vectorised min over the chosen variant's rows; the
    reference's face-by-face loop only runs for samples where a bank update is possible."""
    bank = RefBank(cfg, ref_feat)
    trk = SpanTracker(cfg, fps, total_frames)
    dist = distances if distances is not None else _LiveDistances(face.engine, table)
    fd_add = float(getattr(cfg, "prescan_fd_add", trk.enter))
    cooldown = int(getattr(cfg, "prescan_add_cooldown_samples", 5))
    qmin = float(cfg.face_quality_min)
    last_add = -10 ** 9
    plain_h, flip_h = feats_host
    lazy = table is not None and getattr(table, "lazy", False)
    if lazy:
        flip_h = table.flip_host
    lookahead = 96      # samples whose faces get their flip pass together once the replay needs one of them
    for sample_idx, idx in enumerate(idxs):
        rec = records[idx]
        active = trk.active
        best = FD_NONE
        skipped = trk.gate_skips()
        nfaces = 0
        if not skipped:
            face._frame_idx += 1
            chosen = rec.up
            if chosen is None:
                face._no_face_streak += 1
                face._rot_cycle += 1
                if active:
                    order = (90, 270)
                else:
                    order = ((90, 270)[face._prescan_rr % 2],)
                    face._prescan_rr += 1
                for deg in order:
                    if rec.hits.get(deg, 0) == 0 or rec.heavy_raw.get(deg, 0) == 0:
                        continue
                    if deg in rec.heavy:
                        chosen = rec.heavy[deg]
                        break
            else:
                face._no_face_streak = 0
                face._last_face_idx = face._frame_idx
                face._rot_cycle = 0
            if chosen is not None:
                rows = chosen.rows
                nfaces = len(rows)
                if lazy and active and not table.flip_ready[rows].all():
                    want = [rows]
                    for j in idxs[sample_idx + 1:sample_idx + lookahead]:
                        rj = records[j]
                        want += [rj.up.rows] if rj.up is not None else [v.rows for v in rj.heavy.values()]
                    if table.ensure_flip(getattr(face, "engine", None), np.concatenate(want)):
                        dist.invalidate()
                fdp, fdf = dist.get(bank)
                fds = (fdf if active else fdp)[rows]
                if (sample_idx - last_add) >= cooldown and bool(((fds <= fd_add) & (chosen.quality >= qmin)).any()):
                    # a bank update is possible: the reference's face-by-face order matters (later faces see the new bank)
                    area = (chosen.box[:, 2] - chosen.box[:, 0]) * (chosen.box[:, 3] - chosen.box[:, 1])
                    for i in sorted(range(nfaces), key=lambda i: (chosen.quality[i], area[i]), reverse=True):
                        row = int(rows[i])
                        fdp, fdf = dist.get(bank)
                        fd = float(fdf[row] if active else fdp[row])
                        best = min(best, fd)
                        q = float(chosen.quality[i])
                        if fd <= fd_add and (sample_idx - last_add) >= cooldown and q >= qmin:
                            vec = (flip_h if active else plain_h)[row]
                            if bank.offer(vec, q) in ("added", "replaced"):
                                last_add = sample_idx
                else:
                    best = min(best, float(fds.min()))
        if log is not None:
            log.append(dict(idx=idx, skip=skipped, best=best, active_before=active, nfaces=nfaces))
        trk.observe(idx, best)
    return trk, bank


def encode_records(records: Dict[int, SampleRecord], idxs: Sequence[int], n_rows: int):
    """Flat form of the superset records for pcb_replay: meta int32 [n, PCB_REPLAY_META], quality f64 [rows], area i64 [rows]."""
    meta = np.zeros((len(idxs), L.REPLAY_META), np.int32)
    meta[:, 0] = meta[:, 6] = meta[:, 8] = -1
    quality = np.zeros(max(n_rows, 1), np.float64)
    area = np.zeros(max(n_rows, 1), np.int64)

    def put(v: _Variant):
        r0, k = int(v.rows[0]), len(v.rows)
        if int(v.rows[-1]) - r0 + 1 != k:
            raise ValueError("variant rows must be contiguous")
        quality[r0:r0 + k] = v.quality
        b = v.box.astype(np.int64)
        area[r0:r0 + k] = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
        return r0, k

    for s_i, idx in enumerate(idxs):
        rec = records[idx]
        m = meta[s_i]
        if rec.up is not None:
            m[0], m[1] = put(rec.up)
        for d, deg in enumerate((90, 270)):
            m[2 + d] = rec.hits.get(deg, 0)
            m[4 + d] = rec.heavy_raw.get(deg, 0)
            if deg in rec.heavy:
                m[6 + 2 * d], m[7 + 2 * d] = put(rec.heavy[deg])
    return meta, quality, area


class _BankView:
    """What a `distances.get(bank)` implementation sees of the native bank: array() / version / len()."""

    def __init__(self, rows: Optional[np.ndarray], version: int):
        self._rows, self.version = rows, version

    def array(self):
        return self._rows

    def __len__(self):
        return 0 if self._rows is None else len(self._rows)


def bank_cfg_of(cfg) -> "L.BankCfg":
    wa, wd, wq = _weights3(cfg)
    return L.BankCfg(cap=max(1, int(getattr(cfg, "prescan_bank_max", 64))), dedup=float(getattr(cfg, "prescan_diversity_dedup_cos", 0.968)),
                     margin=float(getattr(cfg, "prescan_replace_margin", 0.010)), wa=wa, wd=wd, wq=wq)


def replay(records: Dict[int, SampleRecord], table, feats_host, idxs: Sequence[int], fps: int, total_frames: int,
           face: FaceEmbedder, ref_feat, cfg, log: Optional[list] = None, distances=None, native: Optional[bool] = None,
           encoded=None):
    """Replay of the reference's sequential loop over precomputed superset records.  The per-sample loop AND the live
    bank run in libpcb200 (pcb_replay / pcb_bank_offer); distances of the rows still ahead are refreshed on the GPU by
    the library itself (pcb_live_refresh: one short launch per bank change, no Python in between).  The only callback
    left is "flip features missing".  `distances` (tests, no GPU) supplies the distances through a callback instead;
    `native=False` (or PCB_PY_REPLAY=1) runs the pure-Python statement of the same loop (`_replay_python`), which the
    tests hold the native one against."""
    import ctypes as C
    if native is None:
        native = os.environ.get("PCB_PY_REPLAY", "0") != "1"
    if not native:
        return _replay_python(records, table, feats_host, idxs, fps, total_frames, face, ref_feat, cfg, log, distances)
    lib = L.load()
    trk = SpanTracker(cfg, fps, total_frames)
    plain_h, flip_h = feats_host if feats_host is not None else (None, None)
    lazy = table is not None and getattr(table, "lazy", False)
    if lazy:
        flip_h = table.flip_host
    n_rows = int(table.count) if table is not None else len(plain_h)
    idxs = list(idxs)
    meta, quality, area = encoded if encoded is not None else encode_records(records, idxs, n_rows)
    meta = np.ascontiguousarray(meta, np.int32)
    quality = np.ascontiguousarray(quality, np.float64)
    area = np.ascontiguousarray(area, np.int64)
    frame_idx = np.asarray(idxs, np.int64)
    fdp = np.full(max(n_rows, 1), FD_NONE, np.float64)
    fdf = np.full(max(n_rows, 1), FD_NONE, np.float64)
    # bank offers read the face's feature row: from host tables when the caller has them (tests, gloo), else straight from
    # the device tables, a row at a time (the replay skips certain duplicates, so only a handful of rows are ever read)
    dev_feats = (distances is None and plain_h is None and table is not None and n_rows > 0 and getattr(table.plain, "is_cuda", False))
    plain_c = flip_c = None
    if not dev_feats:
        plain_c = np.ascontiguousarray(plain_h, np.float32) if n_rows else np.zeros((1, L.FEAT_DIM), np.float32)
        flip_c = np.ascontiguousarray(flip_h, np.float32) if (n_rows and flip_h is not None and len(flip_h)) else np.zeros((max(n_rows, 1), L.FEAT_DIM), np.float32)
        if lazy and n_rows:
            flip_c = table.flip_host            # filled in place by ensure_flip: the library must see the same memory
            assert flip_c.flags["C_CONTIGUOUS"] and flip_c.dtype == np.float32
    ref_rows = None
    if ref_feat is not None:
        ref_rows = np.ascontiguousarray(np.asarray(ref_feat, np.float32).reshape(-1, L.FEAT_DIM))
    bcfg = bank_cfg_of(cfg)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    nb = lib.pcb_bank_create(C.byref(bcfg), ptr(ref_rows) if ref_rows is not None and len(ref_rows) else None,
                             0 if ref_rows is None else len(ref_rows))
    if not nb:
        raise L.PcbError("pcb_bank_create failed")
    eng = getattr(face, "engine", None)
    use_gpu = distances is None
    failure: List[BaseException] = []        # an exception inside a ctypes callback would be printed and swallowed: keep it, abort, re-raise

    def arm_live():
        with torch.cuda.stream(eng.stream):
            both = torch.cat([table.plain[:n_rows], table.flip[:n_rows]], 0).contiguous()
        eng._check(lib.pcb_live_begin(eng.ctx, both.data_ptr(), 2 * n_rows), "pcb_live_begin")

    def on_refresh(_user, bank_rows, n_bank, _slot, _row_lo):
        try:
            rows = np.ctypeslib.as_array(bank_rows, shape=(n_bank, L.FEAT_DIM)).copy() if n_bank else None
            a, b = distances.get(_BankView(rows, int(lib.pcb_bank_version(nb))))
            if n_rows:
                fdp[:n_rows] = a
                fdf[:n_rows] = b
            return 0
        except BaseException as exc:      # noqa: BLE001
            failure.append(exc)
            return -1

    lookahead = 96

    flip_calls = [0, 0.0]

    def on_flip(_user, s_i):
        import time as _t
        t0 = _t.perf_counter()
        flip_calls[0] += 1
        try:
            want = []
            for m in meta[s_i:s_i + lookahead]:
                for c0 in (0, 6, 8):
                    if m[c0] >= 0:
                        want.append(np.arange(m[c0], m[c0] + m[c0 + 1]))
            if want and table.ensure_flip(eng, np.concatenate(want)):
                if use_gpu:
                    arm_live()
                elif hasattr(distances, "invalidate"):
                    distances.invalidate()
            flip_calls[1] += _t.perf_counter() - t0
            return 0
        except BaseException as exc:      # noqa: BLE001
            failure.append(exc)
            return -1

    import time as _t
    t_parts = [_t.perf_counter()]
    try:
        if use_gpu and n_rows:
            arm_live()
            if os.environ.get("PCB_REPLAY_TIMING", "0") == "1":
                eng.sync()
        t_parts.append(_t.perf_counter())
        rc = L.ReplayCfg(enter=trk.enter, exit_thr=trk.exit, fd_add=float(getattr(cfg, "prescan_fd_add", trk.enter)),
                         quality_min=float(cfg.face_quality_min), total_frames=trk.total, pad=trk.pad, min_len=trk.min_len,
                         exit_cool=trk.exit_cool, stride=trk.stride, cooldown=int(getattr(cfg, "prescan_add_cooldown_samples", 5)),
                         fd9_skip=int(trk._fd9_skip), fd9_grace=trk._fd9_grace, fd9_period=trk._fd9_period)
        st = L.ReplayState(frame_idx=int(face._frame_idx), last_face_idx=int(max(face._last_face_idx, -2 ** 62)),
                           no_face_streak=int(face._no_face_streak), rot_cycle=int(face._rot_cycle), prescan_rr=int(face._prescan_rr),
                           trk_active=0)
        n = len(idxs)
        best = np.zeros(max(n, 1), np.float64)
        skip = np.zeros(max(n, 1), np.uint8)
        act = np.zeros(max(n, 1), np.uint8)
        nf = np.zeros(max(n, 1), np.int32)
        max_spans = n + 2
        spans = np.zeros((max_spans, 2), np.int64)
        n_spans = C.c_int32(0)
        n_refresh = C.c_int64(0)
        ready = table.flip_ready.view(np.uint8) if lazy else None
        cb_r, cb_f = L.REPLAY_REFRESH_CB(on_refresh), L.REPLAY_FLIP_CB(on_flip)
        io = L.ReplayIO(meta=ptr(meta), frame_idx=ptr(frame_idx), n_samples=n, n_rows=n_rows, quality=ptr(quality), area=ptr(area),
                        flip_ready=ptr(ready) if ready is not None else None,
                        feat_plain=None if dev_feats else ptr(plain_c), feat_flip=None if dev_feats else ptr(flip_c),
                        feat_plain_dev=table.plain.data_ptr() if dev_feats else None,
                        feat_flip_dev=table.flip.data_ptr() if dev_feats else None,
                        fd_plain=ptr(fdp), fd_flip=ptr(fdf), refresh=cb_r, need_flip=cb_f, user=None, best_out=ptr(best),
                        skip_out=ptr(skip), active_out=ptr(act), nfaces_out=ptr(nf), spans_out=ptr(spans), max_spans=max_spans,
                        n_spans_out=C.pointer(n_spans), refreshes_out=C.pointer(n_refresh))
        err = lib.pcb_replay(eng.ctx if use_gpu else None, C.byref(rc), nb, C.byref(io), C.byref(st))
        t_parts.append(_t.perf_counter())
        if failure:
            raise failure[0]
        if err:
            msg = lib.pcb_last_error(eng.ctx).decode() if (use_gpu and eng is not None) else ""
            raise L.PcbError(f"pcb_replay failed (code {err}) {msg}")
        n_bank = int(lib.pcb_bank_rows(nb))
        bank = RefBank(cfg)
        if n_bank:
            arr = np.ctypeslib.as_array(lib.pcb_bank_data(nb), shape=(n_bank, L.FEAT_DIM)).copy()
            bank.rows = [r for r in arr]
        bank.version = int(lib.pcb_bank_version(nb))
    finally:
        lib.pcb_bank_destroy(nb)
        if use_gpu and eng is not None:
            eng._bank_token = None          # the library rewrote the device bank behind Engine.set_bank's cache
    face._frame_idx, face._last_face_idx = int(st.frame_idx), int(st.last_face_idx)
    face._no_face_streak, face._rot_cycle, face._prescan_rr = int(st.no_face_streak), int(st.rot_cycle), int(st.prescan_rr)
    trk.spans = [(int(a), int(b)) for a, b in spans[:n_spans.value]]
    trk.active = False          # pcb_replay already closed the open span (gui_app.py:1648-1655)
    trk.distance_refreshes = int(n_refresh.value)
    trk.flip_on_demand = (flip_calls[0], round(1000.0 * flip_calls[1], 2))      # (calls, ms inside them)
    trk.replay_parts_ms = [round(1000.0 * (b - a), 3) for a, b in zip(t_parts, t_parts[1:])]      # [arm live table, pcb_replay]
    if log is not None:
        for i, idx in enumerate(idxs):
            log.append(dict(idx=idx, skip=bool(skip[i]), best=float(best[i]), active_before=bool(act[i]), nfaces=int(nf[i])))
    return trk, bank


def _predict_flip_rows(records: Dict[int, SampleRecord], idxs: Sequence[int], fd_plain: np.ndarray, cfg, fps: int,
                       carry_in: bool, margin: float = 0.12) -> np.ndarray:
    """Rows whose flip-TTA feature the replay will most likely read: samples that follow, within the exit cooldown, a
    sample with a face at plain distance <= enter + margin from the *initial* bank (the bank only grows, so live
    distances are not larger), plus the head of a chunk that starts inside another rank's possible active stretch.
    A prediction only: the replay computes whatever is missing on demand, so spans never depend on it."""
    enter = float(cfg.prescan_fd_enter)
    stride = max(1, int(cfg.prescan_stride))
    exit_cool = int(round(max(0.0, float(getattr(cfg, "prescan_exit_cooldown_sec", 0.5))) * fps))
    tail = (exit_cool + stride - 1) // stride + 1
    # ... widened by the span padding on both sides: boundary refinement probes the padded edges in flip-TTA mode
    pad_s = (int(round(max(0.0, float(getattr(cfg, "prescan_pad_sec", 0.0))) * fps)) + stride - 1) // stride + 1
    per = [_rows_of(records[idx]) for idx in idxs]
    hot = [bool(rws) and min(float(fd_plain[r].min()) for r in rws) <= enter + margin for rws in per]
    n = len(per)
    rows: List[np.ndarray] = []
    for k in range(n):
        if any(hot[max(k - tail - pad_s, 0):min(k + pad_s + 1, n)]) or (carry_in and k < tail):
            rows += per[k]
    return np.concatenate(rows) if rows else np.zeros((0,), np.int64)


def _predict_flip_rows_meta(meta: np.ndarray, fd_plain: np.ndarray, cfg, fps: int, carry_in: bool, margin: float = 0.12) -> np.ndarray:
    """`_predict_flip_rows` over the flat sample records (pcb_replay's meta), vectorised: the per-sample Python loop cost
    2-3 ms per 512 samples inside the timed step."""
    n = len(meta)
    if n == 0:
        return np.zeros((0,), np.int64)
    thr = float(cfg.prescan_fd_enter) + margin
    stride = max(1, int(cfg.prescan_stride))
    exit_cool = int(round(max(0.0, float(getattr(cfg, "prescan_exit_cooldown_sec", 0.5))) * fps))
    tail = (exit_cool + stride - 1) // stride + 1
    hot = np.zeros(n, bool)
    for c0 in (0, 6, 8):
        st, cnt = meta[:, c0].astype(np.int64), meta[:, c0 + 1].astype(np.int64)
        has = np.nonzero((st >= 0) & (cnt > 0))[0]
        if len(has):
            # rows of a variant are contiguous: minimum of fd over [start, start + count)
            idx = np.repeat(st[has], cnt[has]) + (np.arange(int(cnt[has].sum())) - np.repeat(np.cumsum(cnt[has]) - cnt[has], cnt[has]))
            mins = np.minimum.reduceat(fd_plain[idx], np.cumsum(cnt[has]) - cnt[has])
            hot[has] |= mins <= thr
    c = np.concatenate([[0], np.cumsum(hot)])                     # c[k] = hot samples among the first k
    # a hot sample among the `tail` samples before this one (the span is then active), widened by the padding on both sides:
    # boundary refinement probes the padded edges of every span in flip-TTA mode and answers them from this table
    pad_s = (int(round(max(0.0, float(getattr(cfg, "prescan_pad_sec", 0.0))) * fps)) + stride - 1) // stride + 1
    k = np.arange(n)
    sel = (c[np.minimum(k + pad_s + 1, n)] - c[np.maximum(k - tail - pad_s, 0)]) > 0
    if carry_in:
        sel[:tail] = True
    rows = []
    for c0 in (0, 6, 8):
        st, cnt = meta[:, c0].astype(np.int64), meta[:, c0 + 1].astype(np.int64)
        pick = np.nonzero(sel & (st >= 0) & (cnt > 0))[0]
        if len(pick):
            rows.append(np.repeat(st[pick], cnt[pick]) + (np.arange(int(cnt[pick].sum())) - np.repeat(np.cumsum(cnt[pick]) - cnt[pick], cnt[pick])))
    return np.concatenate(rows) if rows else np.zeros((0,), np.int64)


class _ShardedTable:
    """The all-gathered face table of a multi-rank pre-scan.  Flip features that no rank predicted are computed by the rank
    that owns the face (it holds the chip) and exchanged; every rank replays identically, so all ranks reach `ensure_flip`
    with the same rows and the exchange is a matched collective."""
    lazy = True

    def __init__(self, local: FaceTable, counts: List[int], rank: int, group, plain, flip, ready, flip_host):
        self.local, self.counts, self.rank, self.group = local, counts, rank, group
        self.base = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self.count = int(self.base[-1])
        self.plain, self.flip, self.flip_ready, self.flip_host = plain, flip, ready, flip_host
        self.flip_passes = 0

    def ensure_flip(self, eng, rows: np.ndarray) -> bool:
        import torch.distributed as dist
        if not len(rows):
            return False
        need = np.unique(np.asarray(rows)[~self.flip_ready[rows]])
        if not len(need):
            return False
        world = len(self.counts)
        # every rank derives the same `need`, so it also knows how many rows each owner contributes: one all_gather of the
        # padded feature blocks is the whole exchange (no pickled objects, no size negotiation)
        owned = [need[(need >= self.base[r]) & (need < self.base[r + 1])] for r in range(world)]
        mine = owned[self.rank] - self.base[self.rank]
        self.local.ensure_flip(eng, mine)
        cap = max(max(len(o) for o in owned), 1)
        cuda = self.flip.is_cuda
        send = torch.zeros((cap, L.FEAT_DIM), dtype=torch.float32, device=self.flip.device)
        if len(mine):
            if cuda:
                with torch.cuda.stream(eng.stream):
                    send[:len(mine)] = self.local.flip.index_select(0, torch.as_tensor(mine, device=self.flip.device))
                torch.cuda.current_stream().wait_stream(eng.stream)
            else:
                send[:len(mine)] = torch.from_numpy(self.local.flip_host[mine])
        recv = torch.empty((world, cap, L.FEAT_DIM), dtype=torch.float32, device=self.flip.device)
        dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.group)
        if cuda:
            eng.stream.wait_stream(torch.cuda.current_stream())
        host = recv.cpu().numpy() if self.flip_host is not None else None
        for r in range(world):
            rows_g = owned[r]
            if not len(rows_g):
                continue
            if self.flip_host is not None:
                self.flip_host[rows_g] = host[r, :len(rows_g)]
            self.flip_ready[rows_g] = True
            if cuda:
                with torch.cuda.stream(eng.stream):
                    self.flip.index_copy_(0, torch.as_tensor(rows_g, device=self.flip.device), recv[r, :len(rows_g)])
            else:
                self.flip[torch.as_tensor(rows_g)] = recv[r, :len(rows_g)]
        return True


def prescan_batched(clip, fps: int, face: FaceEmbedder, ref_feat, cfg, batch: int = 32, log: Optional[list] = None,
                    dist_group=None, stats: Optional[dict] = None, single_rank: bool = False):
    """Throughput pre-scan.  With torch.distributed initialised (dist_group or the default group), the
    sample list is split into contiguous chunks per rank, per-face records are all-gathered and every
    rank replays the same sequence.

    ArcFace work matches the reference's: e(x) for every face; e(flip x) only for faces evaluated while a span is active
    (face_embedder.py:1295).  Which faces those are is a property of the sequential replay, so each rank first embeds the
    flips its own chunk is *predicted* to need (in parallel, large batches) and the replay fills any gap on demand."""
    import torch.distributed as dist
    total = clip.total_frames
    stride = max(1, int(cfg.prescan_stride))
    idxs = sample_indices(total, stride)
    world, rank = 1, 0
    if dist.is_available() and dist.is_initialized() and not single_rank:     # single_rank: this process scans the whole clip alone
        world, rank = dist.get_world_size(dist_group), dist.get_rank(dist_group)
    if int(getattr(cfg, "prescan_probe_imgsz", 512)) > int(face.fast_no_face_imgsz):
        # the upright detector size would then depend on the no-face streak (face_embedder.py:2193-2194), i.e. on the
        # sequential state: the superset is not state independent for such a configuration.  Same results,
        # one extract per sample; with several ranks every rank runs it (no collective is needed).
        return prescan_sequential(clip, fps, face, ref_feat, cfg, log=log)
    eng = face.engine
    import time as _time
    tmark = [("start", _time.perf_counter())]
    # NVTX ranges per phase (PCB_NVTX=1): what a profiler's timeline shows as superset / predicted_flips / gather / replay / refine
    _PHASES = ("superset", "predicted_flips", "gather", "replay", "refine")
    nvtx = torch.cuda.nvtx if (os.environ.get("PCB_NVTX", "0") == "1" and torch.cuda.is_available()) else None
    if nvtx is not None:
        nvtx.range_push("prescan:" + _PHASES[0])

    def mark(name):
        if nvtx is not None:
            nvtx.range_pop()
            k = _PHASES.index(name) + 1
            if k < len(_PHASES):
                nvtx.range_push("prescan:" + _PHASES[k])
        if stats is not None:
            eng.sync()
            tmark.append((name, _time.perf_counter()))

    with _PrescanFaceMode(face, cfg):
        per = (len(idxs) + world - 1) // world
        mine = idxs[rank * per:(rank + 1) * per]
        lazy = os.environ.get("PCB_EAGER_FLIP", "0") != "1"
        # early flip passes + small ArcFace runs: on by default when frames come from the host (the SMs then wait for PCIe
        # between frame batches); PCB_EARLY_FLIP=1 / 0 forces it on / off
        early = None
        mode = os.environ.get("PCB_EARLY_FLIP", "auto")
        if lazy and mode != "0" and (mode == "1" or bool(getattr(clip, "host_resident", False))):
            bank0 = RefBank(cfg, ref_feat)
            if len(bank0):
                aux = face.aux_engine()
                aux.set_bank(bank0.array())      # FaceTable.flush then gets the distances to the initial bank for free
                early = EarlyFlips(aux, cfg, fps, carry_in=rank > 0)
        records, table = compute_superset(clip, mine, face, cfg, batch=batch, lazy_flip=lazy, early=early)
        TIMELINE.mark("superset_end", eng.stream)
        eng.sync()
        if stats is not None:
            stats["early_flip_rows"] = int(table.flip_ready.sum()) if (early is not None and table.count) else 0
        mark("superset")
        if lazy and table.count:
            bank0 = RefBank(cfg, ref_feat)
            if len(bank0):
                eng.set_bank(bank0.array())
                _, s0, _ = eng.match(table.plain, None, None, table.count)
                eng.sync()
                fd0 = 1.0 - s0[:table.count].cpu().numpy().astype(np.float64)
                table.ensure_flip(eng, _predict_flip_rows_meta(table.encoded[0], fd0, cfg, fps, carry_in=rank > 0))
        TIMELINE.mark("predicted_flips_end", eng.stream)
        mark("predicted_flips")
        # host copies of the features: only for the pure-Python replay, a non-NCCL group, or when asked for
        # (PCB_REPLAY_HOST_FEATS=1); the native replay reads the few rows it offers to the bank from the device tables
        want_host = (os.environ.get("PCB_PY_REPLAY", "0") == "1" or os.environ.get("PCB_REPLAY_HOST_FEATS", "0") == "1"
                     or not table.count or not table.plain.is_cuda or (world > 1 and dist.get_backend(dist_group) != "nccl"))
        plain_h = flip_h = None
        if want_host:
            plain_h = table.plain[:table.count].cpu().numpy() if table.count else np.zeros((0, L.FEAT_DIM), np.float32)
            flip_h = (table.flip[:table.count].cpu().numpy() if (table.count and not lazy) else np.zeros((0, L.FEAT_DIM), np.float32))
        local_table = table
        encoded = table.encoded       # flat records built chunk by chunk during the GPU stage
        if world > 1:
            table, plain_h, flip_h, encoded = _gather_shards(eng, encoded, table, plain_h, flip_h, world, dist_group, want_host=want_host)
            records = None
        mark("gather")
        trk, bank = replay(records, table, (plain_h, flip_h), idxs, fps, total, face, ref_feat, cfg, log, encoded=encoded)
        mark("replay")
        if stats is not None:
            stats["bank_rows"], stats["bank_versions"] = len(bank), bank.version
            stats["distance_refreshes"] = getattr(trk, "distance_refreshes", None)
            stats["flip_on_demand"] = getattr(trk, "flip_on_demand", None)
            stats["replay_parts_ms"] = getattr(trk, "replay_parts_ms", None)
        _count_passes(stats, local_table)      # this rank's faces / ArcFace image passes
        spans = trk.finish()
        wmax = int(getattr(cfg, "prescan_max_width", 0))
        shard = None
        if world > 1:
            first_of = [idxs[min(r * per, len(idxs) - 1)] for r in range(world)]
            shard = dict(world=world, rank=rank, group=dist_group,
                         owner=lambda j: max(r for r in range(world) if first_of[r] <= j or r == 0))
        known = None
        if os.environ.get("PCB_REFINE_REUSE", "1") != "0" and table.count and getattr(table.plain, "is_cuda", False):
            known = dict(pos={int(j): k for k, j in enumerate(idxs)}, meta=encoded[0], table=table, dist=_LiveDistances(eng, table))
        spans = _post_process(spans, clip, fps, face, bank, ref_feat, cfg, trk, wmax, batched=batch, stats=stats, shard=shard, known=known)
        mark("refine")
    TIMELINE.dump()
    if stats is not None:
        stats["phase_ms"] = {b[0]: round(1000.0 * (b[1] - a[1]), 2) for a, b in zip(tmark, tmark[1:])}
    out_bank = bank.array()
    return spans, (out_bank if out_bank is not None else ref_feat)


_PINNED: Dict[tuple, torch.Tensor] = {}


def _pinned(tag: str, shape, dtype) -> torch.Tensor:
    """Reusable pinned host buffer (cudaHostAlloc costs ~1 ms per call; the gather runs every pre-scan on every rank)."""
    n = 1
    for d in shape:
        n *= int(d)
    key = (tag, dtype)
    buf = _PINNED.get(key)
    if buf is None or buf.numel() < n:
        buf = torch.empty((max(n, 1),), dtype=dtype).pin_memory()
        _PINNED[key] = buf
    return buf[:n].view(*shape)


def _gather_shards(eng, enc_local, table, plain_h, flip_h, world, group, want_host: bool = True):
    """All-gather of everything the replicated replay needs from the other ranks, as TWO tensor collectives: a 2-word
    header (face rows, samples) and one packed byte buffer per rank

        [ meta int32 [S,10] | quality f64 [F] | area i64 [F] | flip_ready u8 [F] | plain f32 [F,512] | flip f32 [F,512] ]

    padded to the largest rank.  On NCCL the features never leave the device before the collective; the small records
    ride in the same buffer.  -> (merged table, plain_h, flip_h, (meta, quality, area) of all ranks in sample order)."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    lazy = bool(getattr(table, "lazy", False))
    meta_l, q_l, a_l = enc_local
    count = int(table.count)
    n_s = int(len(meta_l))
    backend_cuda = dist.get_backend(group) == "nccl"
    dev = eng.tdev if (backend_cuda and eng is not None) else torch.device("cpu")
    head = torch.tensor([count, n_s], dtype=torch.int64, device=dev)
    heads = torch.empty((world, 2), dtype=torch.int64, device=dev)
    if backend_cuda:
        torch.cuda.current_stream().wait_stream(eng.stream)
    dist.all_gather_into_tensor(heads.view(-1), head, group=group)
    heads_h = heads.cpu().numpy()
    counts = [int(c) for c in heads_h[:, 0]]
    n_samples = [int(c) for c in heads_h[:, 1]]
    cap, scap = max(max(counts), 1), max(max(n_samples), 1)
    r16 = lambda x: (x + 15) // 16 * 16
    o_meta = 0
    o_q = r16(o_meta + scap * L.REPLAY_META * 4)
    o_a = r16(o_q + cap * 8)
    o_r = r16(o_a + cap * 8)
    o_p = r16(o_r + cap)
    o_f = o_p + cap * L.FEAT_DIM * 4
    nbytes = o_f + cap * L.FEAT_DIM * 4
    small = np.zeros(o_p, np.uint8)
    small[o_meta:o_meta + n_s * L.REPLAY_META * 4] = np.ascontiguousarray(meta_l, np.int32).view(np.uint8).reshape(-1)
    small[o_q:o_q + count * 8] = np.ascontiguousarray(q_l[:count], np.float64).view(np.uint8)
    small[o_a:o_a + count * 8] = np.ascontiguousarray(a_l[:count], np.int64).view(np.uint8)
    ready_l = np.asarray(getattr(table, "flip_ready", np.ones(count, bool))[:count], bool) if count else np.zeros((0,), bool)
    small[o_r:o_r + count] = ready_l.view(np.uint8)
    send = torch.zeros((nbytes,), dtype=torch.uint8, device=dev)
    fview = lambda buf, off, n: buf[off:off + n * L.FEAT_DIM * 4].view(torch.float32).view(n, L.FEAT_DIM)
    if backend_cuda:
        with torch.cuda.stream(eng.stream):
            stage = _pinned("gather_small_out", (o_p,), torch.uint8)
            stage.copy_(torch.from_numpy(small))
            send[:o_p].copy_(stage, non_blocking=True)
            if count:
                fview(send, o_p, count).copy_(table.plain[:count])
                fview(send, o_f, count).copy_(table.flip[:count])
        torch.cuda.current_stream().wait_stream(eng.stream)
    else:
        send[:o_p] = torch.from_numpy(small)
        if count:
            fview(send, o_p, count).copy_(torch.from_numpy(np.ascontiguousarray(plain_h[:count], np.float32)))
            flip_src = table.flip_host if lazy else flip_h
            fview(send, o_f, count).copy_(torch.from_numpy(np.ascontiguousarray(flip_src[:count], np.float32)))
    recv = torch.empty((world, nbytes), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv.view(-1), send, group=group)
    base = sum(counts)
    plain_all = torch.cat([fview(recv[r], o_p, counts[r]) for r in range(world)], 0) if base else torch.zeros((1, L.FEAT_DIM), device=dev)
    flip_all = torch.cat([fview(recv[r], o_f, counts[r]) for r in range(world)], 0) if base else torch.zeros((1, L.FEAT_DIM), device=dev)
    if backend_cuda:
        # host copies (the replay hands bank offers a host vector): pinned, one copy per array
        small_h = _pinned("gather_small_in", (world, o_p), torch.uint8)
        small_h.copy_(recv[:, :o_p], non_blocking=True)
        plain_host = flip_host = None
        if want_host:
            # (8 ranks x all features through one host memory system: 5 ms of a 90 ms step on the 8-GPU box -- hence optional)
            ph = _pinned("gather_plain", (max(base, 1), L.FEAT_DIM), torch.float32)
            fh = _pinned("gather_flip", (max(base, 1), L.FEAT_DIM), torch.float32)
            if base:
                ph[:base].copy_(plain_all, non_blocking=True)
                fh[:base].copy_(flip_all, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        eng.stream.wait_stream(torch.cuda.current_stream())
        plain_all, flip_all = plain_all.contiguous(), flip_all.contiguous()
        # views of the reusable pinned buffers: valid until the next gather, i.e. for the rest of this pre-scan (the replay
        # reads them, on-demand flips write flip_host in place)
        small_np = small_h.numpy()
        if want_host:
            plain_host, flip_host = ph[:base].numpy(), fh[:base].numpy()
    else:
        small_np = recv[:, :o_p].numpy()
        plain_host, flip_host = plain_all[:base].numpy().copy(), flip_all[:base].numpy().copy()
        if eng is not None:
            plain_all, flip_all = plain_all.to(eng.tdev).contiguous(), flip_all.to(eng.tdev).contiguous()
        else:
            flip_all = flip_all.clone()
    metas, quals, areas, readies = [], [], [], []
    row0 = 0
    for r in range(world):
        row = np.ascontiguousarray(small_np[r])
        m = row[o_meta:o_meta + n_samples[r] * L.REPLAY_META * 4].view(np.int32).reshape(-1, L.REPLAY_META).copy()
        for c0 in (0, 6, 8):
            m[:, c0] = np.where(m[:, c0] >= 0, m[:, c0] + row0, -1)
        metas.append(m)
        quals.append(row[o_q:o_q + counts[r] * 8].view(np.float64).copy())
        areas.append(row[o_a:o_a + counts[r] * 8].view(np.int64).copy())
        readies.append(row[o_r:o_r + counts[r]].astype(bool))
        row0 += counts[r]
    if lazy:
        ready = np.concatenate(readies) if base else np.zeros((0,), bool)
        new = _ShardedTable(table, counts, rank, group, plain_all, flip_all, ready,
                            np.ascontiguousarray(flip_host, np.float32) if flip_host is not None else None)
    else:
        new = FaceTable()
        new.count = base
        new.plain, new.flip = plain_all, flip_all
    pad1 = lambda xs, dt: np.concatenate(xs).astype(dt) if base else np.zeros(1, dt)
    encoded = (np.concatenate(metas, 0), pad1(quals, np.float64), pad1(areas, np.int64))
    return new, plain_host, flip_host, encoded


# --------------------------------------------------------------------------------------------
# prescan cache (gui_app.py:787-920): same key derivation, file name and array layout
# --------------------------------------------------------------------------------------------
CACHE_KEYS = (
    "prescan_stride", "prescan_max_width", "prescan_decode_max_w", "prescan_face_conf", "prescan_fd_enter", "prescan_fd_add",
    "prescan_fd_exit", "prescan_add_cooldown_samples", "prescan_rot_probe_period", "prescan_probe_imgsz",
    "prescan_no_upscale_det", "prescan_probe_conf", "prescan_heavy_90", "prescan_heavy_180", "prescan_min_segment_sec",
    "prescan_pad_sec", "prescan_bridge_gap_sec", "prescan_exit_cooldown_sec", "prescan_boundary_refine_sec",
    "prescan_refine_stride_min", "prescan_trim_pad", "prescan_skip_trailing_refine", "prescan_refine_budget_sec",
    "prescan_bank_max", "prescan_diversity_dedup_cos", "prescan_replace_margin", "prescan_fd9_skip", "prescan_fd9_grace",
    "prescan_fd9_probe_period", "prescan_weights", "face_model", "clip_face_backbone", "clip_face_pretrained", "use_arcface",
)


def _identity_of(path: str) -> dict:
    p = str(path or "").strip()
    if not p:
        return {"path": "", "missing": True}
    ap = os.path.abspath(p)
    try:
        st = os.stat(ap)
    except Exception:
        return {"path": ap, "missing": True}
    return {"path": ap, "size": int(st.st_size or 0), "mtime_ns": int(st.st_mtime_ns)}


def _plain(v):
    if isinstance(v, (tuple, list)):
        return [_plain(x) for x in v]
    if isinstance(v, (np.floating, np.integer)):
        return v.item()
    return v


def cache_meta(cfg, fps: float, total_frames: int) -> dict:
    meta = {
        "version": 1,
        "video": _identity_of(getattr(cfg, "video", "")),
        "refs": [_identity_of(p) for p in (q.strip() for q in str(getattr(cfg, "ref", "") or "").split(";")) if p],
        "fps": round(float(fps or 0.0), 6),
        "total_frames": int(total_frames or 0),
        "settings": {k: _plain(getattr(cfg, k, None)) for k in CACHE_KEYS},
    }
    meta["key"] = hashlib.sha256(json.dumps(meta, sort_keys=True, separators=(",", ":")).encode("utf-8")).hexdigest()
    return meta


def cache_file(cfg, meta: dict, root=None) -> Path:
    base = Path(root) if root is not None else Path(str(getattr(cfg, "prescan_cache_dir", "prescan_cache") or "prescan_cache"))
    return base / f"{meta.get('key') or ''}.npz"


def save_cache(cfg, fps: float, total_frames: int, spans, ref_face_feat, root=None) -> Optional[Path]:
    mode = str(getattr(cfg, "prescan_cache_mode", "auto") or "auto").lower()
    if mode not in ("auto", "refresh", "reuse"):
        return None
    meta = cache_meta(cfg, fps, total_frames)
    path = cache_file(cfg, meta, root)
    path.parent.mkdir(parents=True, exist_ok=True)
    if ref_face_feat is None:
        feat, has = np.zeros((0, 0), np.float32), np.array([0], np.uint8)
    else:
        feat = np.asarray(ref_face_feat, np.float32)
        feat = feat.reshape(1, -1) if feat.ndim == 1 else feat
        has = np.array([1], np.uint8)
    tmp = path.with_suffix(path.suffix + ".tmp")
    with open(tmp, "wb") as fh:
        np.savez_compressed(fh, meta=np.array(json.dumps(meta, sort_keys=True), dtype=np.str_),
                            spans=np.asarray(spans or [], np.int64).reshape(-1, 2), ref_face_feat=feat, has_ref=has)
    os.replace(tmp, path)
    return path


def load_cache(cfg, fps: float, total_frames: int, root=None):
    """-> (hit, spans, ref_feat, meta)"""
    mode = str(getattr(cfg, "prescan_cache_mode", "auto") or "auto").lower()
    if mode not in ("auto", "reuse"):
        return False, [], None, None
    meta = cache_meta(cfg, fps, total_frames)
    path = cache_file(cfg, meta, root)
    if not path.is_file():
        return False, [], None, meta
    try:
        with np.load(str(path), allow_pickle=False) as z:
            stored = json.loads(str(z["meta"].item()))
            if stored.get("key") != meta["key"] or stored.get("version") != meta["version"]:
                return False, [], None, meta
            arr = np.asarray(z["spans"], np.int64).reshape(-1, 2)
            spans = [(int(s), int(e)) for s, e in arr.tolist() if e >= s]
            has = bool(int(np.asarray(z["has_ref"]).reshape(-1)[0])) if "has_ref" in z.files else False
            ref = None
            if has and "ref_face_feat" in z.files:
                a = np.asarray(z["ref_face_feat"], np.float32)
                if a.size:
                    ref = a.reshape(a.shape[0], -1) if a.ndim >= 2 else a.reshape(1, -1)
            return True, spans, ref, meta
    except Exception:
        return False, [], None, meta
