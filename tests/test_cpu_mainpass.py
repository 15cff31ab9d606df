"""CPU: host logic of the main-pass identity driver (person_capture_b200/mainpass.py) against the oracle restatement
(oracle/mainpass.py) on scripted detections -- lock-face ROI geometry, miss counter, cadence, fallback, segment jumps,
runtime bank learning.  No GPU: a scripted face source stands in for FaceEmbedder on both sides."""
import numpy as np
import pytest

from oracle import mainpass as OM
from oracle import prescan as OP
from person_capture_b200 import mainpass as MP
from person_capture_b200.params import PrescanParams


def unit(v):
    v = np.asarray(v, np.float32)
    return v / np.linalg.norm(v)


class ScriptedFace:
    """extract(img) looks the frame index up from pixel (0,0) of the FULL frame; ROI crops carry a marker in their own
    (0,0) pixel only if the crop starts at the frame origin, so crops are resolved through `self.current`."""

    def __init__(self, script, W, H):
        self.script, self.W, self.H = script, W, H
        self.current = None
        self.calls = []

    def extract(self, img, *, imgsz=None):
        h, w = img.shape[:2]
        full = (h == self.H and w == self.W)
        if full:
            self.current = int(img[0, 0, 0]) + 256 * int(img[0, 0, 1])
        self.calls.append((self.current, h, w, imgsz))
        faces = []
        for (x1, y1, x2, y2, q, feat) in self.script.get(self.current, []):
            if full:
                faces.append(dict(bbox=np.array([x1, y1, x2, y2], np.int32), quality=q, feat=feat))
            else:
                ox, oy = self.roi_origin
                if x1 >= ox and y1 >= oy and x2 <= ox + w and y2 <= oy + h:
                    faces.append(dict(bbox=np.array([x1 - ox, y1 - oy, x2 - ox, y2 - oy], np.int32), quality=q, feat=feat))
        faces.sort(key=lambda f: (f["quality"], (f["bbox"][2] - f["bbox"][0]) * (f["bbox"][3] - f["bbox"][1])), reverse=True)
        return faces


class _Clip:
    def __init__(self, n, W, H, face):
        self.total_frames, self.W, self.H, self.face = n, W, H, face

    def host(self, i):
        a = np.zeros((self.H, self.W, 3), np.uint8)
        a[0, 0, 0], a[0, 0, 1] = i % 256, i // 256
        return _Tracked(a, self.face)


class _Tracked(np.ndarray):
    """ndarray that remembers the origin of the last slice taken from it (the ROI the driver cuts)."""

    def __new__(cls, arr, face):
        obj = np.asarray(arr).view(cls)
        obj._face = face
        return obj

    def __getitem__(self, key):
        out = super().__getitem__(key)
        if isinstance(key, tuple) and len(key) == 2 and all(isinstance(k, slice) for k in key):
            self._face.roi_origin = (key[1].start or 0, key[0].start or 0)
        return np.asarray(out)


def _scenario(seed, n, W, H):
    rng = np.random.default_rng(seed)
    target = unit(rng.normal(size=512))
    script = {}
    x, y = 200.0, 120.0
    for i in range(n):
        faces = []
        present = (i // 23) % 3 != 2
        x = float(np.clip(x + rng.normal(0, 6), 10, W - 90))
        y = float(np.clip(y + rng.normal(0, 4), 10, H - 90))
        if present and rng.random() < 0.85:
            s = int(rng.integers(40, 80))
            f = unit(target + rng.normal(0, rng.uniform(0.02, 0.07), 512))
            faces.append((int(x), int(y), int(x) + s, int(y) + s, float(rng.uniform(30, 400)), f))
        for _ in range(int(rng.integers(0, 3))):
            s = int(rng.integers(20, 70))
            ox, oy = int(rng.integers(0, W - s)), int(rng.integers(0, H - s))
            faces.append((ox, oy, ox + s, oy + s, float(rng.uniform(30, 400)), unit(rng.normal(size=512))))
        script[i] = faces
    return target, script


@pytest.mark.parametrize("seed,stride,cadence,learn,use_qv", [(0, 1, 12, False, True), (1, 2, 5, True, True), (2, 3, 0, False, False),
                                                              (3, 1, 7, True, False)])
def test_main_pass_matches_oracle(seed, stride, cadence, learn, use_qv):
    n, W, H = 240, 640, 360
    target, script = _scenario(seed, n, W, H)
    cfg = PrescanParams(frame_stride=stride, face_fullframe_cadence=cadence, learn_bank_runtime=learn, face_visible_uses_quality=use_qv,
                        face_thresh=0.45, face_quality_min=70.0, prescan_fd_add=0.3, prescan_add_cooldown_samples=3,
                        lock_face_roi_max_misses=2, prescan_bank_max=4)
    ref = unit(target + np.random.default_rng(99).normal(0, 0.03, 512))[None]
    spans = [(5, 90), (120, 200), (230, 239)]

    of = ScriptedFace(script, W, H)
    oclip = _Clip(n, W, H, of)
    olog = []
    ohits = OM.main_pass(lambda i: oclip.host(i) if i < n else None, 24.0, n, spans, of, ref, cfg, log=olog)

    gf = ScriptedFace(script, W, H)
    gclip = _Clip(n, W, H, gf)
    glog = []
    fd_fn = lambda faces, bank: np.array([OP.fd_min(f["feat"], bank) for f in faces])
    ghits = MP.main_pass(gclip, 24.0, spans, gf, ref, cfg, log=glog, device_frames=False, fd_fn=fd_fn)

    assert [(r["idx"], r["site"], r["roi"], r["accept"]) for r in glog] == [(r["idx"], r["site"], r["roi"], r["accept"]) for r in olog]
    assert [(h["idx"], h["site"], h["face_box"]) for h in ghits] == [(h["idx"], h["site"], h["face_box"]) for h in ohits]
    assert all(abs(a["fd"] - b["fd"]) < 1e-6 for a, b in zip(ghits, ohits))
    assert gf.calls == of.calls                      # same extract calls (frames, ROI sizes, imgsz) in the same order
    sites = {h["site"] for h in ohits}
    assert "lock_roi" in sites and ({"fullframe", "fallback"} & sites)
    assert any(r["roi"] is not None and not r["accept"] for r in olog)     # lock-ROI misses happen


def test_expand_xyxy_matches_reference_rule():
    for box, px, py, W, H in [((10.2, 20.7, 50.1, 60.9), 16.0, 20.5, 100, 80), ((0, 0, 5, 5), 16, 16, 64, 64), ((90, 70, 99, 79), 30, 30, 100, 80)]:
        assert MP.expand_xyxy(box, px, py, W, H) == OM.expand_xyxy(box, px, py, W, H)


@pytest.mark.parametrize("seed,use_qv", [(0, True), (1, False), (2, True), (3, True)])
def test_person_site_and_arbitration_match_oracle(seed, use_qv):
    """Per-person-crop site (boxes given) + frame-level arbitration / lock gate (App. C rules 2-4): the product statement and
    the oracle's take the same extract calls, keep the same candidates and choose the same one, frame after frame, with
    the lock state (hits, previous box) carried along."""
    n, W, H = 160, 640, 360
    target, script = _scenario(seed, n, W, H)
    rng = np.random.default_rng(100 + seed)
    cfg = PrescanParams(face_thresh=0.45, face_quality_min=70.0, face_visible_uses_quality=use_qv)
    cfg.face_margin_min, cfg.score_margin, cfg.lock_face_thresh = 0.05, 0.03, 0.28
    ref = unit(target + np.random.default_rng(99).normal(0, 0.03, 512))[None]
    of, gf = ScriptedFace(script, W, H), ScriptedFace(script, W, H)
    oclip, gclip = _Clip(n, W, H, of), _Clip(n, W, H, gf)
    fd_fn = lambda faces, bank: np.array([OP.fd_min(f["feat"], bank) for f in faces])
    st = MP.MainPassIdentity(gf, ref, cfg, fd_fn=fd_fn)
    state = {k: dict(hits=0, prev=None) for k in ("o", "g")}
    chosen_n = dropped = gated = retried = 0
    for i in range(n):
        # person boxes: around every scripted face (sometimes too tight, so that only the padded retry finds the face), plus an empty one
        boxes = []
        for (x1, y1, x2, y2, q, f) in script[i]:
            if rng.random() < 0.25:
                boxes.append((x1 + 3, y1 + 3, x2 + 40, y2 + 60))        # cuts the face: first extract is empty
            else:
                boxes.append((max(0, x1 - 20), max(0, y1 - 15), min(W, x2 + 25), min(H, y2 + 80)))
        boxes.append((5, 5, 45, 85))
        fo, fg = oclip.host(i), gclip.host(i)
        of.extract(fo); gf.extract(fg)                 # registers the frame index with the scripted sources
        n_calls = len(of.calls)
        oc, oinfo = OM.person_crop_candidates(fo, boxes, of, ref, cfg)
        retried += int(len(of.calls) - n_calls > len(boxes))
        gc, gvis = MP.person_site(st, fg, boxes)
        assert gf.calls == of.calls
        assert gvis == oinfo["any_face_visible"]
        assert [(c["i"], c["box"], c["face_box"]) for c in gc] == [(c["i"], c["box"], c["face_box"]) for c in oc], i
        assert all(abs(a["fd"] - b["fd"]) < 1e-6 and a["quality"] == b["quality"] for a, b in zip(gc, oc))
        cd = 3 if i % 50 == 0 else 0
        och = OM.arbitrate(oc, oinfo["any_face_visible"], cfg, lock_hits=state["o"]["hits"], locked_face=state["o"]["prev"] is not None,
                           prev_box=state["o"]["prev"], seek_cooldown=cd)
        gch = MP.arbitrate(gc, gvis, cfg, lock_hits=state["g"]["hits"], locked_face=state["g"]["prev"] is not None,
                           prev_box=state["g"]["prev"], seek_cooldown=cd)
        assert (och is None) == (gch is None), i
        if och is not None:
            assert och["i"] == gch["i"]
            chosen_n += 1
            for k, ch in (("o", och), ("g", gch)):
                state[k]["hits"] += 1
                state[k]["prev"] = ch["box"]
        dropped += int(bool(oc) and och is None)
        gated += int(len(oc) < sum(1 for b in boxes[:-1]))
    assert chosen_n >= 20 and retried >= 5 and gated >= 10


def test_arbitration_rules():
    cfg = PrescanParams()
    c = lambda i, fd, box=(0, 0, 10, 10): dict(i=i, box=box, fd=fd, score=fd, quality=100.0, face_box=box, area=100, sharp=0.0)
    for arb in (OM.arbitrate, MP.arbitrate):
        assert arb([], True, cfg) is None
        assert arb([c(0, 0.20), c(1, 0.23)], True, cfg) is None                       # ambiguous: fd2 - fd1 < face_margin_min
        assert arb([c(0, 0.20), c(1, 0.23)], False, cfg)["i"] == 0                    # no visible face: margin rule off, score margin keeps the best
        assert arb([c(0, 0.30), c(1, 0.20), c(2, 0.40)], True, cfg)["i"] == 1
        # lock gate: the closer candidate fails the IoU gate, the other passes lock_face_thresh
        far, near = c(0, 0.10, (500, 300, 560, 360)), c(1, 0.25, (12, 12, 60, 60))
        assert arb([far, near], True, cfg, lock_hits=1, locked_face=True, prev_box=(10, 10, 60, 60))["i"] == 1
        assert arb([far, near], True, cfg, lock_hits=1, locked_face=True, prev_box=(10, 10, 60, 60), seek_cooldown=2)["i"] == 0
        assert arb([far, near], True, cfg, lock_hits=0, locked_face=True, prev_box=(10, 10, 60, 60))["i"] == 0
        weak = c(1, 0.40, (12, 12, 60, 60))                                           # overlaps but fd > lock_face_thresh: fall back to the best
        assert arb([far, weak], True, cfg, lock_hits=1, locked_face=True, prev_box=(10, 10, 60, 60))["i"] == 0


class StatefulScriptedFace(ScriptedFace):
    """ScriptedFace with the FaceEmbedder counters that make the main pass sequential across spans: extract() counts calls,
    tracks the no-face streak / last face, and on EMPTY frames reveals the frame's "rotated" faces only when the adaptive
    rotation rule fires (recent hit, or the absolute call counter hits the rot_every_n grid) -- as the real class does."""
    rot_phase, rot_every_n, rot_after_hit_frames = 3, 5, 2

    def __init__(self, script, rotated, W, H):
        super().__init__(script, W, H)
        self.rotated = rotated
        self._frame_idx, self._no_face_streak, self._last_face_idx, self._rot_cycle = 0, 0, -10 ** 9, 0
        self.rot_consults = None

    def extract(self, img, *, imgsz=None):
        self._frame_idx += 1
        faces = super().extract(img, imgsz=imgsz)
        if faces:
            self._no_face_streak, self._last_face_idx, self._rot_cycle = 0, self._frame_idx, 0
            return faces
        self._no_face_streak += 1
        recent = (self._frame_idx - self._last_face_idx) <= self.rot_after_hit_frames
        if not recent and self.rot_consults is not None:
            self.rot_consults.append(self._frame_idx)
        if not (recent or (self._frame_idx + self.rot_phase) % self.rot_every_n == 0):
            return []
        h, w = img.shape[:2]
        if not (h == self.H and w == self.W):
            return []
        self._rot_cycle += 1
        out = [dict(bbox=np.array([x1, y1, x2, y2], np.int32), quality=q, feat=feat) for (x1, y1, x2, y2, q, feat) in self.rotated.get(self.current, [])]
        return out


def _sharded_case(seed):
    n, W, H = 400, 640, 360
    target, script = _scenario(seed, n, W, H)
    rng = np.random.default_rng(500 + seed)
    rotated = {}
    for i in range(n):
        if not script[i] or rng.random() < 0.3:
            script[i] = [] if rng.random() < 0.5 else script[i]
        if not script[i]:                      # empty upright frame: the target may be there rotated
            s = int(rng.integers(40, 80))
            rotated[i] = [(100, 80, 100 + s, 80 + s, 200.0, unit(target + rng.normal(0, 0.04, 512)))]
    cfg = PrescanParams(frame_stride=1, face_fullframe_cadence=4, face_thresh=0.45, face_quality_min=70.0, lock_face_roi_max_misses=3)
    ref = unit(target + np.random.default_rng(99).normal(0, 0.03, 512))[None]
    spans = [(3, 40), (47, 60), (66, 120), (131, 150), (158, 199), (205, 260), (262, 263), (270, 330), (338, 399)]
    return n, W, H, script, rotated, cfg, ref, spans


def _sharded_worker(rank, world, port, seed, q):
    import os
    import sys
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here)); sys.path.insert(0, here)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, W, H, script, rotated, cfg, ref, spans = _sharded_case(seed)
    face = StatefulScriptedFace(script, rotated, W, H)
    clip = _Clip(n, W, H, face)
    fd_fn = lambda faces, bank: np.array([OP.fd_min(f["feat"], bank) for f in faces])
    stats = {}
    hits = MP.main_pass_sharded(clip, 24.0, spans, face, ref, cfg, device_frames=False, fd_fn=fd_fn, stats=stats)
    q.put((rank, [(h["idx"], h["site"], h["face_box"], round(h["fd"], 9)) for h in hits], stats["rounds"], stats["frames_processed"]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,seed", [(2, 0), (3, 1), (4, 2)])
def test_sharded_main_pass_equals_sequential(world, seed):
    """Spans split into contiguous blocks over `world` gloo ranks: every rank ends with exactly the hits of the sequential
    main pass, although lock box, miss counter and the embedder's call counter (adaptive rotation grid) cross the block
    boundaries; the fix-up re-runs only a fraction of the frames."""
    import os
    import torch.multiprocessing as mp
    n, W, H, script, rotated, cfg, ref, spans = _sharded_case(seed)
    face = StatefulScriptedFace(script, rotated, W, H)
    clip = _Clip(n, W, H, face)
    fd_fn = lambda faces, bank: np.array([OP.fd_min(f["feat"], bank) for f in faces])
    log = []
    want = [(h["idx"], h["site"], h["face_box"], round(h["fd"], 9)) for h in
            MP.main_pass(clip, 24.0, spans, face, ref, cfg, device_frames=False, fd_fn=fd_fn, log=log)]
    assert len(want) >= 60 and {"lock_roi", "fullframe", "fallback"} <= {w[1] for w in want}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30100 + (os.getpid() % 1500) + world
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r, hits, rounds, frames in outs:
        assert hits == want, (r, [a for a, b in zip(hits, want) if a != b][:3])
        assert rounds <= world + 1
    assert sum(o[3] for o in outs) == len(log)        # after the fix-up every processed frame is accounted for exactly once


def test_index_rows_columns():
    rows = MP.index_rows([dict(idx=48, site="fullframe", fd=0.31, quality=120.0, face_box=(10, 20, 60, 90))], 24.0)
    assert MP.INDEX_COLUMNS[:9] == ["frame", "time_secs", "score", "face_dist", "reid_dist", "x1", "y1", "x2", "y2"]
    assert rows == [[48, 2.0, 0.31, 0.31, "", 10, 20, 60, 90, "", "", ""]]
