"""ORACLE (test infrastructure): CPU restatement of the reference pre-scan.

Follows person_capture/gui_app.py:
  _fd_min                          :660-674
  _prescan_weights / face conf     :676-706
  _prescan_cache_meta/_path/_load/_save :787-920
  _stream_ref_bank_update          :922-986
  _prescan (seed bank, loop, hysteresis, tail) :1140-1222, :1468-1655
  bridge / _refine_edges / bridge  :1657-1845
  reference-bank build             :4517-4556
UI command handling, decode/seek and Qt status are out of scope: frames come from a
`get_frame(i)` callable.  The wall-clock refine budget (gui_app.py:1692-1696) makes the
reference timing dependent; the oracle runs with the budget disabled.
Pinned against the reference itself: tests/golden/reference_golden.npz holds the outputs of the unmodified
`Processor._fd_min / _stream_ref_bank_update / _prescan / _prescan_cache_*` run in the build container on seeded scripts
(tests/golden/make_reference_golden.py); tests/test_cpu_reference_golden.py demands identical actions, banks, spans, per-call
knob state and cache key / file from this module.
"""
from __future__ import annotations

import ast
import hashlib
import json
import os
from pathlib import Path
from typing import Callable, List, Optional, Tuple

import cv2
import numpy as np

CACHE_KEYS = (
    "prescan_stride", "prescan_max_width", "prescan_decode_max_w", "prescan_face_conf",
    "prescan_fd_enter", "prescan_fd_add", "prescan_fd_exit", "prescan_add_cooldown_samples",
    "prescan_rot_probe_period", "prescan_probe_imgsz", "prescan_no_upscale_det", "prescan_probe_conf",
    "prescan_heavy_90", "prescan_heavy_180", "prescan_min_segment_sec", "prescan_pad_sec",
    "prescan_bridge_gap_sec", "prescan_exit_cooldown_sec", "prescan_boundary_refine_sec",
    "prescan_refine_stride_min", "prescan_trim_pad", "prescan_skip_trailing_refine",
    "prescan_refine_budget_sec", "prescan_bank_max", "prescan_diversity_dedup_cos",
    "prescan_replace_margin", "prescan_fd9_skip", "prescan_fd9_grace", "prescan_fd9_probe_period",
    "prescan_weights", "face_model", "clip_face_backbone", "clip_face_pretrained", "use_arcface",
)


def fd_min(feat, bank) -> float:
    if feat is None or bank is None:
        return 9.0
    v = np.asarray(feat, dtype=np.float32).reshape(-1)
    v = v / max(float(np.linalg.norm(v)), 1e-6)
    b = np.asarray(bank, dtype=np.float32)
    if b.ndim == 1:
        return 1.0 - float(np.dot(v, b))
    if b.size == 0:
        return 9.0
    sims = b @ v
    if sims.size == 0:
        return 9.0
    return 1.0 - float(np.max(sims))


def prescan_weights(cfg):
    w = getattr(cfg, "prescan_weights", (0.70, 0.25, 0.05))
    if isinstance(w, str):
        raw = w.strip()
        if raw:
            try:
                w = json.loads(raw)
            except Exception:
                try:
                    w = ast.literal_eval(raw)
                except Exception:
                    w = (0.70, 0.25, 0.05)
    if not isinstance(w, (list, tuple)) or len(w) < 3:
        return 0.70, 0.25, 0.05
    try:
        return float(w[0]), float(w[1]), float(w[2])
    except Exception:
        return 0.70, 0.25, 0.05


def face_conf_value(cfg) -> float:
    try:
        return min(0.95, max(0.01, float(getattr(cfg, "prescan_face_conf", 0.5))))
    except Exception:
        return 0.5


def bank_update(bank_list: List[np.ndarray], bank, vec_new, quality, cfg):
    """-> (bank array or None, action in {skip, added, dup, replaced}, replaced index)."""
    if vec_new is None:
        return bank, "skip", None
    cap = max(1, int(getattr(cfg, "prescan_bank_max", 64)))
    dedup = float(getattr(cfg, "prescan_diversity_dedup_cos", 0.968))
    margin = float(getattr(cfg, "prescan_replace_margin", 0.010))
    wa, wd, wq = prescan_weights(cfg)
    v = np.asarray(vec_new, dtype=np.float32).reshape(-1)
    n = float(np.linalg.norm(v))
    if n <= 1e-6:
        return bank, "skip", None
    v = v / n
    B = np.asarray(bank if bank is not None else bank_list, dtype=np.float32)
    if B.ndim == 1:
        B = B.reshape(1, -1)
    if B.size == 0:
        bank_list.append(v)
        return np.vstack(bank_list).astype(np.float32), "added", None
    sims = B @ v
    if sims.size > 0 and float(sims.max()) >= dedup:
        return bank, "dup", None
    anchor = B[0]
    ca = max(-1.0, min(1.0, float(np.dot(anchor, v))))
    fd_a = float(np.sqrt(max(0.0, 2.0 - 2.0 * ca)))
    nn = float(sims.max()) if sims.size else 0.0
    qt = float(min(max(quality or 0.0, 0.0), 1000.0) / 300.0)
    s_new = wa * (1.0 - fd_a) + wd * (1.0 - nn) + wq * qt
    if len(bank_list) < cap:
        bank_list.append(v)
        return np.vstack(bank_list).astype(np.float32), "added", None
    G = B @ B.T
    np.fill_diagonal(G, -1.0)
    nn_each = G.max(axis=1)
    ca_each = np.clip(B @ anchor, -1.0, 1.0)
    fd_each = np.sqrt(np.maximum(0.0, 2.0 - 2.0 * ca_each))
    s_bank = wa * (1.0 - fd_each) + wd * (1.0 - nn_each)
    worst = int(np.argmin(s_bank))
    if s_new > float(s_bank[worst]) + margin:
        bank_list[worst] = v
        return np.vstack(bank_list).astype(np.float32), "replaced", worst
    return bank, "skip", None


def build_reference_bank(face, ref_images: List[np.ndarray], cfg):
    """gui_app.py:4517-4556: each image and its horizontal flip, best face, streaming update."""
    bank_list: List[np.ndarray] = []
    bank = None
    for img in ref_images:
        for aug in (img, cv2.flip(img, 1)):
            bf = face.best_face(face.extract(aug))
            if bf and bf.get("feat") is not None:
                bank, _, _ = bank_update(bank_list, bank, bf["feat"], float(bf.get("quality", 0.0)), cfg)
    return np.vstack(bank_list).astype(np.float32) if bank_list else None


def _bridge(spans, gap):
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


def prescan(get_frame: Callable[[int], np.ndarray], fps: int, total_frames: int, face, ref_feat, cfg,
            log: Optional[list] = None):
    """-> (spans [(s, e)], bank).  `fps` is what the caller passes: int(round(fps)) (gui_app.py:5049)."""
    if ref_feat is None:
        bank_list = []
    else:
        arr = np.asarray(ref_feat, dtype=np.float32)
        if arr.ndim == 1:
            arr = arr.reshape(1, -1)
        arr = arr / np.maximum(np.linalg.norm(arr, axis=1, keepdims=True), 1e-6)
        bank_list = [r.copy() for r in arr]
    bank = np.vstack(bank_list).astype(np.float32) if bank_list else None
    stride = max(1, int(cfg.prescan_stride))
    pad = int(round(cfg.prescan_pad_sec * fps))
    min_len = int(round(cfg.prescan_min_segment_sec * fps))
    Wmax = int(getattr(cfg, "prescan_max_width", 0))
    enter, exit_ = float(cfg.prescan_fd_enter), float(cfg.prescan_fd_exit)
    fd_add = float(getattr(cfg, "prescan_fd_add", enter))
    old_conf = getattr(face, "conf", 0.5)
    old_adapt = getattr(face, "rot_adaptive", True)

    def apply_runtime():
        face.conf = face_conf_value(cfg)
        face._probe_conf = float(getattr(cfg, "prescan_probe_conf", 0.03))
        face._prescan_period = int(getattr(cfg, "prescan_rot_probe_period", 3))
        face._prescan_probe_imgsz = int(getattr(cfg, "prescan_probe_imgsz", 512))
        face._prescan_no_upscale_det = bool(getattr(cfg, "prescan_no_upscale_det", True))
        face._high_90 = int(getattr(cfg, "prescan_heavy_90", 1536))
        face._high_180 = int(getattr(cfg, "prescan_heavy_180", 1280))

    apply_runtime()
    try:
        face.configure_rotation_strategy(adaptive=False)
        face.set_prescan_fast(True, mode="rr")
        face.set_prescan_hint(escalate=False)
        apply_runtime()
        cooldown = int(getattr(cfg, "prescan_add_cooldown_samples", 5))
        last_add = -10 ** 9
        spans: List[Tuple[int, int]] = []
        active = False
        start = 0
        neg_run = 0
        processed = 0
        fd9_streak = 0
        i = 0
        while i < total_frames:
            frame = get_frame(i)
            if frame is None:
                break
            idx = i
            sample_idx = processed
            processed += 1
            h, w = frame.shape[:2]
            face._prescan_rr_mode = "full" if active else "rr"
            face.set_prescan_hint(escalate=active)
            best = 9.0
            skip = False
            if (not active) and bool(getattr(cfg, "prescan_fd9_skip", True)):
                grace = max(0, int(getattr(cfg, "prescan_fd9_grace", 1)))
                period = max(1, int(getattr(cfg, "prescan_fd9_probe_period", 2)))
                if fd9_streak >= grace and (fd9_streak % period) != 0:
                    skip = True
            nfaces = 0
            if not skip:
                if w > Wmax:
                    nh = int(round(h * (Wmax / float(w))))
                    frame = cv2.resize(frame, (Wmax, nh), interpolation=cv2.INTER_AREA)
                faces = face.extract(frame)
                nfaces = len(faces)
                for f in faces:
                    feat = f.get("feat")
                    if feat is None:
                        continue
                    fd = fd_min(feat, bank)
                    best = min(best, fd)
                    if fd <= fd_add and (sample_idx - last_add) >= cooldown and f.get("quality", 1e9) >= cfg.face_quality_min:
                        bank, action, _ = bank_update(bank_list, bank, feat, float(f.get("quality", 0.0)), cfg)
                        if action in ("added", "replaced"):
                            last_add = sample_idx
            fd9_streak = fd9_streak + 1 if best >= 8.99 else 0
            if log is not None:
                log.append(dict(idx=idx, skip=skip, best=best, active_before=active, nfaces=nfaces))
            if best <= enter:
                if not active:
                    active = True
                    fd9_streak = 0
                    start = idx
                neg_run = 0
            elif active:
                neg_run += 1
                exit_cool = int(round(max(0.0, float(getattr(cfg, "prescan_exit_cooldown_sec", 0.5))) * fps))
                if neg_run * stride >= exit_cool or best >= exit_:
                    s = max(0, start - pad)
                    e = min(total_frames - 1, idx + pad)
                    if e - s + 1 >= min_len:
                        if spans and s <= spans[-1][1] + 1:
                            spans[-1] = (spans[-1][0], max(spans[-1][1], e))
                        else:
                            spans.append((s, e))
                    active = False
                    neg_run = 0
                    fd9_streak = 0
            i = idx + 1 + min(max(0, stride - 1), max(0, total_frames - idx - 1))
        if active:
            s = max(0, start - pad)
            e = total_frames - 1
            if e - s + 1 >= min_len:
                if spans and s <= spans[-1][1] + 1:
                    spans[-1] = (spans[-1][0], max(spans[-1][1], e))
                else:
                    spans.append((s, e))
        gap = int(round(cfg.prescan_bridge_gap_sec * fps))
        if spans and getattr(cfg, "prescan_bridge_gap_sec", 0) > 0:
            spans = _bridge(spans, gap)
        spans = _refine_edges(spans, get_frame, fps, total_frames, face, bank, ref_feat, cfg, min_len, Wmax, enter)
        if spans and getattr(cfg, "prescan_bridge_gap_sec", 0) > 0:
            spans = _bridge(spans, gap)
    finally:
        face.configure_rotation_strategy(adaptive=bool(old_adapt))
        face.set_prescan_fast(False)
        face.set_prescan_hint(escalate=False)
        face.conf = old_conf
    return spans, (bank if bank is not None else ref_feat)


def _refine_edges(sp_list, get_frame, fps, total_frames, face, bank_live, ref_feat, cfg, min_len, Wmax, enter):
    if not sp_list:
        return sp_list
    stride_ref = max(1, min(int(max(1, cfg.prescan_stride) // 4), int(getattr(cfg, "prescan_refine_stride_min", 3))))
    win = int(round(max(0.0, float(getattr(cfg, "prescan_boundary_refine_sec", 0.75))) * fps))
    pad_frames = int(round(max(0.0, float(cfg.prescan_pad_sec)) * fps))
    search = max(pad_frames, win)
    trim = bool(getattr(cfg, "prescan_trim_pad", True))
    rr_old = getattr(face, "_prescan_rr_mode", "rr")
    face._prescan_rr_mode = "full"
    face.set_prescan_hint(escalate=True)
    bank = bank_live if bank_live is not None else ref_feat

    def matches(j, first_only):
        frame = get_frame(j)
        if frame is None:
            return False
        h, w = frame.shape[:2]
        if Wmax > 0 and w > Wmax:
            sc = float(Wmax) / float(w)
            frame = cv2.resize(frame, (int(round(w * sc)), int(round(h * sc))), interpolation=cv2.INTER_AREA)
        faces = face.extract(frame)
        hit = False
        for f in faces or ():
            feat = f.get("feat")
            if feat is not None and fd_min(feat, bank) <= enter:
                hit = True
                if first_only:
                    break
        return hit

    refined = []
    for s, e in sp_list:
        ls, le = s, e
        skip_right = bool(getattr(cfg, "prescan_skip_trailing_refine", True)) and (e >= total_frames - 1)
        best_left = None
        j = s
        while j <= min(e, s + search):
            if matches(j, True):
                best_left = j
                break
            j += stride_ref
        if best_left is not None and trim:
            ls = max(s, best_left)
        last_good = None
        if not skip_right:
            j = max(ls, e - search)
            while j <= e:
                if matches(j, False):
                    last_good = j
                j += stride_ref
        if last_good is not None and trim:
            le = min(e, last_good)
        if le >= ls and (le - ls + 1) >= min_len:
            refined.append((ls, le))
    face.set_prescan_hint(escalate=False)
    face._prescan_rr_mode = rr_old
    return refined


# ---- cache (gui_app.py:709-735, 787-920) -------------------------------------------------

def _file_identity(path: str) -> dict:
    p = str(path or "").strip()
    if not p:
        return {"path": "", "missing": True}
    ap = os.path.abspath(p)
    try:
        st = os.stat(ap)
        return {"path": ap, "size": int(st.st_size or 0), "mtime_ns": int(st.st_mtime_ns)}
    except Exception:
        return {"path": ap, "missing": True}


def _jsonable(v):
    if isinstance(v, (tuple, list)):
        return [_jsonable(x) for x in v]
    if isinstance(v, (np.floating, np.integer)):
        return v.item()
    return v


def cache_meta(cfg, fps: float, total_frames: int) -> dict:
    settings = {k: _jsonable(getattr(cfg, k, None)) for k in CACHE_KEYS}
    refs = [p.strip() for p in str(getattr(cfg, "ref", "") or "").split(";") if p.strip()]
    meta = {
        "version": 1,
        "video": _file_identity(getattr(cfg, "video", "")),
        "refs": [_file_identity(p) for p in refs],
        "fps": round(float(fps or 0.0), 6),
        "total_frames": int(total_frames or 0),
        "settings": settings,
    }
    key_json = json.dumps(meta, sort_keys=True, separators=(",", ":"))
    meta["key"] = hashlib.sha256(key_json.encode("utf-8")).hexdigest()
    return meta


def cache_path(cfg, meta, root: Path) -> Path:
    return Path(root) / f"{meta.get('key') or ''}.npz"


def save_cache(cfg, fps, total_frames, spans, ref_face_feat, root: Path):
    mode = str(getattr(cfg, "prescan_cache_mode", "auto") or "auto").lower()
    if mode not in ("auto", "refresh", "reuse"):
        return None
    meta = cache_meta(cfg, fps, total_frames)
    path = cache_path(cfg, meta, root)
    path.parent.mkdir(parents=True, exist_ok=True)
    spans_arr = np.asarray(spans or [], dtype=np.int64).reshape(-1, 2)
    if ref_face_feat is None:
        feat_arr = np.zeros((0, 0), dtype=np.float32)
        has_ref = np.array([0], dtype=np.uint8)
    else:
        feat_arr = np.asarray(ref_face_feat, dtype=np.float32)
        if feat_arr.ndim == 1:
            feat_arr = feat_arr.reshape(1, -1)
        has_ref = np.array([1], dtype=np.uint8)
    tmp = path.with_suffix(path.suffix + ".tmp")
    with open(tmp, "wb") as f:
        np.savez_compressed(f, meta=np.array(json.dumps(meta, sort_keys=True), dtype=np.str_),
                            spans=spans_arr, ref_face_feat=feat_arr, has_ref=has_ref)
    os.replace(tmp, path)
    return path


def load_cache(cfg, fps, total_frames, root: Path):
    """-> (hit, spans, ref_feat, meta)"""
    mode = str(getattr(cfg, "prescan_cache_mode", "auto") or "auto").lower()
    if mode not in ("auto", "reuse"):
        return False, [], None, None
    meta = cache_meta(cfg, fps, total_frames)
    path = cache_path(cfg, meta, root)
    if not path.is_file():
        return False, [], None, meta
    try:
        with np.load(str(path), allow_pickle=False) as data:
            stored = json.loads(str(data["meta"].item()))
            if stored.get("key") != meta.get("key") or stored.get("version") != meta.get("version"):
                return False, [], None, meta
            arr = np.asarray(data["spans"], dtype=np.int64).reshape(-1, 2)
            spans = [(int(s), int(e)) for s, e in arr.tolist() if int(e) >= int(s)]
            has = data["has_ref"] if "has_ref" in data.files else np.array([0], dtype=np.uint8)
            ref = None
            if bool(int(np.asarray(has).reshape(-1)[0])) and "ref_face_feat" in data:
                a = np.asarray(data["ref_face_feat"], dtype=np.float32)
                if a.size > 0:
                    ref = a.reshape(a.shape[0], -1) if a.ndim >= 2 else a.reshape(1, -1)
            return True, spans, ref, meta
    except Exception:
        return False, [], None, meta
