"""Scripted per-frame face scenarios shared by the CPU host-logic tests and the reference-golden generator
(tests/golden/make_reference_golden.py): a stand-in for FaceEmbedder whose faces are looked up from prepared records, and the
conversion of such a scenario into the superset records / face table the GPU stage would have produced."""
import numpy as np

from person_capture_b200 import prescan as PS


def unit(v):
    return (v / np.linalg.norm(v)).astype(np.float32)


class FakeFace:
    """Stands in for FaceEmbedder on both sides: faces are looked up from prepared records with the
    reference's fast-pre-scan branch structure (upright | rr/full rotation choice, flip when escalated)."""

    def __init__(self, scenario):
        self.sc = scenario
        self.conf = 0.5
        self.rot_adaptive = True
        self._prescan_rr = 0
        self._prescan_rr_mode = "rr"
        self._prescan_escalate = False
        self._frame_idx = 0
        self._no_face_streak = 0
        self._last_face_idx = -10 ** 9
        self._rot_cycle = 0
        self.fast_no_face_imgsz = 512
        self.engine = None

    def configure_rotation_strategy(self, **kw):
        if kw.get("adaptive") is not None:
            self.rot_adaptive = bool(kw["adaptive"])

    def set_prescan_fast(self, enable, mode="rr"):
        self._fast = enable
        self._prescan_rr_mode = mode
        if enable:
            self._prescan_rr = 0

    def set_prescan_hint(self, escalate=False):
        self._prescan_escalate = bool(escalate)

    def extract(self, frame):
        idx = int(frame[0, 0, 0]) + 256 * int(frame[0, 0, 1])
        rec = self.sc[idx]
        self._frame_idx += 1
        chosen = rec.get("up")
        if chosen is None:
            if self._prescan_rr_mode == "rr":
                order = ((90, 270)[self._prescan_rr % 2],)
                self._prescan_rr += 1
            else:
                order = (90, 270)
            for deg in order:
                if rec.get(("heavy", deg)) is not None:
                    chosen = rec[("heavy", deg)]
                    break
        if chosen is None:
            return []
        out = []
        for box, q, plain, flip in chosen:
            out.append(dict(bbox=np.array(box, np.int32), quality=float(q), feat=flip if self._prescan_escalate else plain))
        out.sort(key=lambda f: (f["quality"], (f["bbox"][2] - f["bbox"][0]) * (f["bbox"][3] - f["bbox"][1])), reverse=True)
        return out


def make_scenario(rng, n, target):
    sc = {}
    for i in range(n):
        rec = {}
        present = (i // 17) % 2 == 1
        def faces(k):
            fs = []
            for _ in range(k):
                if present and rng.random() < 0.8:
                    base = target + rng.normal(0, rng.uniform(0.02, 0.06), 512)
                else:
                    base = rng.normal(size=512)
                plain = unit(base + rng.normal(0, 0.01, 512))
                flip = unit(base + rng.normal(0, 0.01, 512))
                x, y, s = rng.integers(0, 300), rng.integers(0, 200), rng.integers(20, 90)
                fs.append(((x, y, x + s, y + s), float(rng.uniform(20, 400)), plain, flip))
            return fs
        r = rng.random()
        if r < 0.6:
            rec["up"] = faces(int(rng.integers(1, 4)))
        elif r < 0.8:
            for deg in (90, 270):
                if rng.random() < 0.5:
                    rec[("heavy", deg)] = faces(1)
        sc[i] = rec
    return sc


def to_records(sc):
    """Scenario -> the superset records / face table the GPU stage would have produced."""
    records, plains, flips = {}, [], []
    row = 0
    for i, rec in sc.items():
        r = PS.SampleRecord(i)
        def variant(fs):
            nonlocal row
            rows = np.arange(row, row + len(fs))
            row += len(fs)
            for f in fs:
                plains.append(f[2]); flips.append(f[3])
            return PS._Variant(np.array([f[0] for f in fs], np.int32), np.array([f[1] for f in fs], np.float64), rows)
        if rec.get("up") is not None:
            r.up = variant(rec["up"])
        for deg in (90, 270):
            fs = rec.get(("heavy", deg))
            r.hits[deg] = 1 if fs is not None else 0
            if fs is not None:
                r.heavy_raw[deg] = 1
                r.heavy[deg] = variant(fs)
        records[i] = r
    P = np.stack(plains) if plains else np.zeros((0, 512), np.float32)
    Fl = np.stack(flips) if flips else np.zeros((0, 512), np.float32)
    return records, P, Fl
