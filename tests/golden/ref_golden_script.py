"""Inputs of the reference-golden vectors (tests/golden/reference_golden.npz), shared by the generator
(make_reference_golden.py, which feeds them to the unmodified reference) and by tests/test_cpu_reference_golden.py (which
feeds them to the oracle and to the product's host logic).  Everything here is seeded and cheap; nothing imports the
reference."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from scenario_helpers import FakeFace, make_scenario, unit  # noqa: E402
from person_capture_b200 import synth  # noqa: E402

ACTIONS = ["skip", "added", "dup", "replaced"]
EXTRACT_SCRFD = "scrfd_2.5g_bnkps"


# ------------------------------------------------------------------------------------------------- units
def unit_cases():
    """[(face crop BGR uint8, 5 landmarks float32 crop-local)]: canonical, permuted, mirrored, rolled, tied, degenerate."""
    rng = np.random.default_rng(20260)
    cases = []
    base = np.array([[0.32, 0.38], [0.68, 0.37], [0.50, 0.58], [0.36, 0.76], [0.65, 0.77]], np.float64)
    for k in range(18):
        h, w = int(rng.integers(28, 150)), int(rng.integers(28, 150))
        crop = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 1.2)
        pts = base * [w, h] + rng.normal(0, 0.02 * min(h, w), (5, 2))
        mode = k % 9
        if mode == 1:
            pts = pts[rng.permutation(5)]
        elif mode == 2:                                   # rolled by a large angle about the crop centre (eye-roll branch)
            a = np.deg2rad(float(rng.choice([-70, -35, 25, 50, 110])))
            c, s = np.cos(a), np.sin(a)
            ctr = np.array([w / 2.0, h / 2.0])
            pts = (pts - ctr) @ np.array([[c, -s], [s, c]]).T + ctr
        elif mode == 3:                                   # eyes tie on x: canon refuses
            pts[1, 0] = pts[0, 0]
        elif mode == 4:                                   # nose above the eyes: canon refuses
            pts[2, 1] = pts[:2, 1].min() - 3.0
        elif mode == 5:                                   # small roll (< 8 degrees): plain resize branch
            pts[1, 1] = pts[0, 1] + 0.05 * (pts[1, 0] - pts[0, 0])
            pts[2, 1] = pts[:2, 1].min() - 2.0            # and canon refuses so the roll path is the one used
        elif mode == 6:                                   # upside down
            pts[:, 1] = h - 1 - pts[:, 1]
        elif mode == 7:                                   # mouth corners tie on x
            pts[4, 0] = pts[3, 0]
        elif mode == 8:                                   # collapsed landmarks
            pts[:] = pts[0]
        cases.append((np.ascontiguousarray(crop), pts.astype(np.float32)))
    return cases


# ------------------------------------------------------------------------------------------------- extract
_CLIPS = {}


def _clip(seed):
    if seed not in _CLIPS:
        _CLIPS[seed] = synth.ClipSpec(416, 234, 400, seed=seed, target_segments=[(0, 399)], distractor_prob=0.6)
    return _CLIPS[seed]


def frame_of(key):
    """('face', seed, idx, rot) | ('blank', seed, rot) | ('roi', seed, idx, side) | ('big', seed, idx)"""
    kind = key[0]
    rots = {0: None, 90: cv2.ROTATE_90_CLOCKWISE, 180: cv2.ROTATE_180, 270: cv2.ROTATE_90_COUNTERCLOCKWISE}
    if kind == "face":
        f = _clip(key[1]).frame(key[2])
        return f if not key[3] else np.ascontiguousarray(cv2.rotate(f, rots[key[3]]))
    if kind == "blank":
        f = synth.background(np.random.default_rng(key[1]), 234, 416, clutter=4)
        return f if not key[2] else np.ascontiguousarray(cv2.rotate(f, rots[key[2]]))
    if kind == "roi":                                     # a small crop around the target (main-pass lock ROI, up-scaled by imgsz)
        f, truth = _clip(key[1]).frame_with_truth(key[2])
        x1, y1, x2, y2 = [int(v) for v in truth[0][1]]
        s = key[3]
        cx, cy = (x1 + x2) // 2, (y1 + y2) // 2
        return np.ascontiguousarray(f[max(0, cy - s):cy + s, max(0, cx - s):cx + s])
    if kind == "big":
        return synth.ClipSpec(1280, 720, 60, seed=key[1], target_segments=[(0, 59)]).frame(key[2])
    raise KeyError(key)


def extract_script():
    """Knob changes and extract() calls, in order.  Mirrors how the reference drives its embedder: default (main-pass) mode,
    the pre-scan's fast mode with rr / full rotation cadence and flip escalation, explicit imgsz (ROI / 4K sites)."""
    ops = []
    ex = lambda key, imgsz=None: ops.append(("extract", key, imgsz))
    # --- main-pass mode: adaptive rotations, streak-dependent size
    for i in (11, 12):
        ex(("face", 1001, i, 0))
    for k in range(14):
        ex(("blank", 50 + k, 0))
    ex(("face", 1001, 40, 90))
    ex(("face", 1001, 41, 90))
    ex(("face", 1001, 42, 0))
    ex(("roi", 1001, 60, 40), 320)
    ex(("roi", 1001, 61, 56), 640)
    ex(("big", 1001, 5), 1280)
    # --- pre-scan mode as Processor._prescan sets it (gui_app.py:1162-1196)
    ops.append(("attr", "conf", 0.5))
    ops.append(("attr", "_probe_conf", 0.03))
    ops.append(("attr", "_prescan_period", 3))
    ops.append(("attr", "_prescan_probe_imgsz", 384))
    ops.append(("attr", "_prescan_no_upscale_det", True))
    ops.append(("attr", "_high_90", 1536))
    ops.append(("attr", "_high_180", 1280))
    ops.append(("knob", "configure_rotation_strategy", dict(adaptive=False)))
    ops.append(("knob", "set_prescan_fast", dict(enable=True, mode="rr")))
    ops.append(("knob", "set_prescan_hint", dict(escalate=False)))
    for i in (20, 21):
        ex(("face", 1001, i, 0))
    for k in range(7):
        ex(("blank", 80 + k, 0))
    for rot in (90, 270, 180, 90, 270, 90):
        ex(("face", 1001, 100 + rot // 90, rot))
    # active span: full rotations + flip
    ops.append(("attr", "_prescan_rr_mode", "full"))
    ops.append(("knob", "set_prescan_hint", dict(escalate=True)))
    ex(("face", 1001, 30, 0))
    ex(("face", 1001, 31, 270))
    ex(("face", 1001, 32, 90))
    ex(("face", 1001, 33, 180))
    ex(("blank", 95, 0))
    ex(("blank", 96, 0))
    # back to idle, different probe settings
    ops.append(("attr", "_prescan_rr_mode", "rr"))
    ops.append(("knob", "set_prescan_hint", dict(escalate=False)))
    ops.append(("attr", "_prescan_period", 2))
    ops.append(("attr", "_prescan_probe_imgsz", 512))
    ops.append(("attr", "_prescan_no_upscale_det", False))
    ex(("face", 1001, 34, 0))
    for k in range(5):
        ex(("blank", 110 + k, 0))
    ex(("face", 1001, 35, 270))
    ex(("face", 1001, 36, 90))
    # --- pre-scan ends (gui_app.py:1851-1862)
    ops.append(("knob", "configure_rotation_strategy", dict(adaptive=True)))
    ops.append(("knob", "set_prescan_fast", dict(enable=False)))
    ops.append(("knob", "set_prescan_hint", dict(escalate=False)))
    ex(("face", 1001, 37, 0))
    ex(("blank", 120, 0))
    return ops


# ------------------------------------------------------------------------------------------------- bank
def unit_vec(seed):
    return unit(np.random.default_rng(900 + seed).normal(size=512))


def bank_cases():
    """name -> (cfg overrides, offered vectors, qualities, seed rows)"""
    cases = {}
    rng = np.random.default_rng(31)
    base = unit(rng.normal(size=512))
    feats = [unit(base + rng.normal(0, float(rng.choice([0.005, 0.05, 0.3])), 512)) * np.float32(rng.uniform(0.5, 2.0)) for _ in range(160)]
    feats[17] = np.zeros(512, np.float32)                                   # zero vector: skip
    quals = [float(rng.uniform(0, 1200)) for _ in range(160)]
    cases["cap4"] = (dict(prescan_bank_max=4), feats, quals, [unit(base + rng.normal(0, 0.02, 512))])
    rng = np.random.default_rng(32)
    feats2 = [unit(rng.normal(size=512) + 0.8 * base * np.sqrt(512)) for _ in range(120)]
    cases["cap8_w"] = (dict(prescan_bank_max=8, prescan_weights="[0.5, 0.4, 0.1]", prescan_replace_margin=0.0,
                            prescan_diversity_dedup_cos=0.9), feats2, [float(rng.uniform(0, 500)) for _ in range(120)], [])
    return cases


# ------------------------------------------------------------------------------------------------- prescan
class PipeLikeCap:
    """Frame source with the interface Processor._prescan / _seek_to use on the reference's ffmpeg pipe reader
    (`_is_hdr_pipe`: seeks are exact frame seeks, gui_app.py:3993-4045).  Frames are constant images whose colour encodes
    the index (blue = idx % 256, green = idx // 256), so an INTER_AREA downscale leaves the code intact."""
    _is_hdr_pipe = True

    def __init__(self, n, h=2, w=2):
        self.n, self.h, self.w = int(n), int(h), int(w)
        self.pos = 0
        self.cur = -1
        self.reads = []

    def _next_frame_index(self):
        return self.pos

    def frame(self, i):
        if i < 0 or i >= self.n:
            return None
        a = np.zeros((self.h, self.w, 3), np.uint8)
        a[..., 0], a[..., 1] = i % 256, i // 256
        return a

    def get(self, prop):
        if prop == cv2.CAP_PROP_POS_FRAMES:
            return float(self.pos)
        if prop == cv2.CAP_PROP_FRAME_COUNT:
            return float(self.n)
        return 0.0

    def set(self, prop, val):
        if prop == cv2.CAP_PROP_POS_FRAMES:
            self.pos = int(val)
            return True
        return False

    def grab(self):
        if self.pos >= self.n:
            return False
        self.cur = self.pos
        self.pos += 1
        return True

    def retrieve(self):
        f = self.frame(self.cur)
        self.reads.append(self.cur)
        return (f is not None), f

    def read(self):
        if not self.grab():
            return False, None
        return self.retrieve()

    def release(self):
        pass


class RecordingFakeFace(FakeFace):
    """FakeFace that also records what the pre-scan loop handed to it at every extract call."""

    def __init__(self, scenario):
        super().__init__(scenario)
        self.calls = []

    def extract(self, frame):
        idx = int(frame[0, 0, 0]) + 256 * int(frame[0, 0, 1])
        self.calls.append((idx, int(self._prescan_rr_mode == "full"), int(bool(self._prescan_escalate)), int(frame.shape[1])))
        return super().extract(frame)


_BASE_CFG = dict(prescan_max_width=10 ** 6, prescan_fd_add=0.3, prescan_add_cooldown_samples=2, face_quality_min=50.0,
                 prescan_min_segment_sec=0.25, prescan_pad_sec=0.1, prescan_exit_cooldown_sec=0.2, prescan_boundary_refine_sec=0.0,
                 prescan_refine_budget_sec=0.0)          # no wall-clock cap on the refine pass: the vectors must not depend on timing


def prescan_cases():
    """name -> dict(cfg overrides, fps, total_frames, scenario seed, ref kind, frame size)"""
    c = {}
    c["s1"] = dict(cfg=dict(_BASE_CFG, prescan_stride=1), fps=24, n=240, seed=0)
    c["s3_refine"] = dict(cfg=dict(_BASE_CFG, prescan_stride=3, prescan_boundary_refine_sec=0.75, prescan_pad_sec=0.25), fps=24, n=240, seed=1)
    c["s2_cap3"] = dict(cfg=dict(_BASE_CFG, prescan_stride=2, prescan_bank_max=3, prescan_add_cooldown_samples=0), fps=24, n=240, seed=2)
    c["s6_fd9"] = dict(cfg=dict(_BASE_CFG, prescan_stride=6, prescan_fd9_grace=2, prescan_fd9_probe_period=3, prescan_bridge_gap_sec=1.0,
                                prescan_min_segment_sec=0.4, prescan_boundary_refine_sec=0.5, prescan_refine_stride_min=2), fps=30, n=480, seed=5)
    c["noref"] = dict(cfg=dict(_BASE_CFG, prescan_stride=4), fps=24, n=120, seed=6, ref=None)
    c["s4_resize"] = dict(cfg=dict(_BASE_CFG, prescan_stride=4, prescan_max_width=4, prescan_boundary_refine_sec=0.3), fps=24, n=240, seed=7, w=8, h=6)
    c["trail_skip"] = dict(cfg=dict(_BASE_CFG, prescan_stride=2, prescan_boundary_refine_sec=0.4, prescan_skip_trailing_refine=True), fps=24, n=238, seed=8)
    c["trail_refine"] = dict(cfg=dict(_BASE_CFG, prescan_stride=2, prescan_boundary_refine_sec=0.4, prescan_skip_trailing_refine=False,
                                      prescan_trim_pad=True), fps=24, n=238, seed=8)
    c["no_fd9_no_bridge"] = dict(cfg=dict(_BASE_CFG, prescan_stride=5, prescan_fd9_skip=False, prescan_bridge_gap_sec=0.0, prescan_trim_pad=False,
                                          prescan_boundary_refine_sec=0.5), fps=25, n=300, seed=9)
    c["two_row_ref"] = dict(cfg=dict(_BASE_CFG, prescan_stride=3, prescan_fd_enter=0.40, prescan_fd_exit=0.60, prescan_fd_add=0.2), fps=24, n=240, seed=10,
                            ref="two")
    return c


def prescan_inputs(case):
    """-> (scenario, ref_feat or None) for a prescan case (seeded)."""
    rng = np.random.default_rng(case["seed"])
    target = unit(rng.normal(size=512))
    sc = make_scenario(rng, case["n"], target)
    kind = case.get("ref", "one")
    if kind is None:
        ref = None
    elif kind == "two":
        ref = np.stack([unit(target + rng.normal(0, 0.03, 512)), unit(target + rng.normal(0, 0.05, 512))]) * np.float32(1.7)   # un-normalised on purpose
    else:
        ref = unit(target + rng.normal(0, 0.03, 512))[None]
    return sc, ref


class FrameCap(PipeLikeCap):
    """PipeLikeCap over real frames."""

    def __init__(self, frames):
        super().__init__(len(frames))
        self.frames = frames

    def frame(self, i):
        return self.frames[i] if 0 <= i < self.n else None


FULL_CFG = dict(face_model="scrfd_2.5g_bnkps", prescan_stride=3, prescan_max_width=416, prescan_fd_enter=0.25, prescan_fd_exit=0.32,
                prescan_fd_add=0.12, face_quality_min=40.0, prescan_min_segment_sec=0.5, prescan_pad_sec=0.25, prescan_bridge_gap_sec=0.25,
                prescan_exit_cooldown_sec=0.25, prescan_boundary_refine_sec=0.5, prescan_add_cooldown_samples=2, prescan_probe_imgsz=384,
                prescan_refine_budget_sec=0.0)     # 0 disables the reference's wall-clock cap on the refine pass (timing dependent)
FULL_N, FULL_FPS, FULL_SEED = 144, 24, 1001


def full_clip_frames():
    clip = synth.ClipSpec(640, 360, FULL_N, seed=FULL_SEED)
    return [clip.frame(i) for i in range(FULL_N)], synth.reference_image(1, 512, seed=FULL_SEED)


# ------------------------------------------------------------------------------------------------- curator
class CannedFace:
    """extract() returns canned faces chosen by the image shape (and records the canvases it was given)."""

    def __init__(self):
        self.seen = []

    def extract(self, bgr):
        import zlib
        self.seen.append((bgr.shape[0], bgr.shape[1], zlib.crc32(np.ascontiguousarray(bgr).tobytes()) & 0xFFFFFFFF))
        rng = np.random.default_rng(int(bgr[:8, :8].sum()) + bgr.shape[0] * 7 + bgr.shape[1])
        k = int(rng.integers(0, 4))
        faces = []
        for _ in range(k):
            x, y = int(rng.integers(0, bgr.shape[1] - 40)), int(rng.integers(0, bgr.shape[0] - 40))
            sd = int(rng.integers(16, 40))
            faces.append(dict(bbox=np.array([x, y, x + sd, y + sd], np.int32), quality=float(rng.choice([50.0, 120.0, 120.0, 300.0])),
                              feat=unit(rng.normal(size=512)) * np.float32(rng.uniform(0.5, 2))))
        return faces


def curator_images():
    rng = np.random.default_rng(77)
    shapes = [(300, 200), (200, 300), (640, 640), (90, 1000), (1000, 90), (480, 854), (64, 64), (721, 333)]
    return [cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 1.0) for h, w in shapes]


# ------------------------------------------------------------------------------------------------- main-pass geometry
def geometry_cases():
    """[(box xyxy float, pad_x, pad_y, W, H)] incl. boxes at / beyond the frame edges, and box pairs for IoU."""
    rng = np.random.default_rng(404)
    cases = []
    for _ in range(60):
        W, Hh = int(rng.integers(64, 4000)), int(rng.integers(64, 2200))
        x1, y1 = float(rng.uniform(-40, W)), float(rng.uniform(-40, Hh))
        w, h = float(rng.uniform(0.5, 600)), float(rng.uniform(0.5, 600))
        cases.append(((x1, y1, x1 + w, y1 + h), float(rng.uniform(0, 200)), float(rng.uniform(0, 200)), W, Hh))
    cases.append(((0.0, 0.0, 10.0, 10.0), 16.0, 16.0, 100, 100))
    cases.append(((95.5, 95.5, 130.0, 130.0), 3.25, 0.0, 100, 100))
    cases.append(((-50.0, -50.0, -10.0, -10.0), 1.0, 1.0, 64, 64))
    pairs = [(tuple(float(v) for v in rng.uniform(0, 300, 4)), tuple(float(v) for v in rng.uniform(0, 300, 4))) for _ in range(40)]
    pairs = [((min(a[0], a[2]), min(a[1], a[3]), max(a[0], a[2]), max(a[1], a[3])), (min(b[0], b[2]), min(b[1], b[3]), max(b[0], b[2]), max(b[1], b[3])))
             for a, b in pairs]
    pairs.append(((0.0, 0.0, 10.0, 10.0), (10.0, 10.0, 20.0, 20.0)))
    pairs.append(((5.0, 5.0, 5.0, 5.0), (0.0, 0.0, 10.0, 10.0)))
    return cases, pairs


# ------------------------------------------------------------------------------------------------- reference x ONNX graphs
RO_SCRFD, RO_ARC = "scrfd_2.5g_bnkps", "arcface_r50"
RO_FRAMES = [(416, 234, 77, i) for i in (0, 9, 18, 27, 36, 45, 54, 63)]          # (W, H, clip seed, frame): test_extract_matches_oracle's clip


def ro_frame(key):
    W, Hh, seed, i = key
    return synth.ClipSpec(W, Hh, 120, seed=seed).frame(i)


# the bench's models on the bench's frame shape (tests/test_gpu_headline.py::test_extract_960x540_r100_matches_oracle)
RH_SCRFD, RH_ARC = "scrfd_10g_bnkps", "arcface_r100"
RH_FRAME_IDS = list(range(4, 120, 11))


def rh_frame(i):
    clip = _CLIPS.setdefault(("rh", 1002), synth.ClipSpec(1920, 1080, 120, seed=1002, distractor_prob=1.0))
    return cv2.resize(clip.frame(i), (960, 540), interpolation=cv2.INTER_AREA)


# the pre-scan of tests/test_gpu_e2e.py::test_prescan_spans_match_oracle (seed 1006, stride 3), run by the reference on ONNX graphs
RP_CFG = dict(face_model="scrfd_2.5g_bnkps", prescan_stride=3, prescan_max_width=416, prescan_fd_enter=0.62, prescan_fd_exit=0.72,
              prescan_fd_add=0.50, face_quality_min=40.0, prescan_min_segment_sec=0.5, prescan_pad_sec=0.25, prescan_bridge_gap_sec=0.25,
              prescan_exit_cooldown_sec=0.25, prescan_boundary_refine_sec=0.5, prescan_refine_budget_sec=0.0)
RP_SEED, RP_N, RP_FPS = 1006, 144, 24


def rp_clip_frames():
    clip = synth.ClipSpec(640, 360, RP_N, seed=RP_SEED)
    return [clip.frame(i) for i in range(RP_N)], synth.reference_image(1, 512, seed=RP_SEED)


# the headline pre-scan of tests/test_gpu_headline.py (SCRFD-10G + iResNet-100, 1080p -> 960 wide, stride 2)
RQ_CFG = dict(face_model="scrfd_10g_bnkps", prescan_stride=2, prescan_max_width=960, prescan_fd_enter=0.62, prescan_fd_exit=0.72,
              prescan_fd_add=0.50, face_quality_min=40.0, prescan_min_segment_sec=0.5, prescan_pad_sec=0.25, prescan_bridge_gap_sec=0.25,
              prescan_exit_cooldown_sec=0.25, prescan_boundary_refine_sec=0.5, prescan_refine_budget_sec=0.0)
RQ_SEED, RQ_N, RQ_FPS = 2001, 72, 24


def rq_clip_frames():
    clip = synth.ClipSpec(1920, 1080, RQ_N, seed=RQ_SEED)
    return [clip.frame(i) for i in range(RQ_N)], synth.reference_image(1, 512, seed=RQ_SEED)


CACHE_DIR = "/tmp/pcb_reference_golden_cache"
CACHE_CFG = dict(prescan_stride=5, prescan_max_width=512, prescan_fd_enter=0.41, prescan_weights=(0.6, 0.3, 0.1), face_model="scrfd_10g_bnkps")


def cache_files():
    """Creates the two files the cache key fingerprints (fixed path, size and mtime) -> (video path, ref string)."""
    os.makedirs(CACHE_DIR, exist_ok=True)
    video = os.path.join(CACHE_DIR, "clip.mp4")
    refs = [os.path.join(CACHE_DIR, "ref_a.png"), os.path.join(CACHE_DIR, "ref_b.png")]
    with open(video, "wb") as f:
        f.write(b"\x00\x01video-bytes" * 100)
    with open(refs[0], "wb") as f:
        f.write(b"ref-bytes" * 50)
    if os.path.exists(refs[1]):
        os.remove(refs[1])                                       # second reference is missing on purpose
    os.utime(video, ns=(1_700_000_000_000_000_000, 1_700_000_000_123_456_000))
    os.utime(refs[0], ns=(1_700_000_100_000_000_000, 1_700_000_100_987_654_000))
    return video, ";".join(refs)
