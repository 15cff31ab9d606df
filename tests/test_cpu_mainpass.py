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
