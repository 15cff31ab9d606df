"""CPU: the product's host logic (RefBank, SpanTracker, replay of superset records, cache layout) against the
oracle's restatement of the reference loop, on randomly generated per-sample records (no GPU, no networks)."""
import json

import numpy as np
import pytest

from oracle import prescan as OP
from person_capture_b200 import prescan as PS
from person_capture_b200.params import PrescanParams


from scenario_helpers import FakeFace, make_scenario, to_records, unit  # noqa: E402


class NumpyDistances:
    def __init__(self, plain, flip):
        self.p, self.f = plain, flip

    def get(self, bank):
        B = bank.array()
        if B is None:
            return np.full(len(self.p), 9.0), np.full(len(self.f), 9.0)
        return (np.array([OP.fd_min(v, B) for v in self.p]), np.array([OP.fd_min(v, B) for v in self.f]))


@pytest.mark.parametrize("seed,stride,bank_max", [(0, 1, 64), (1, 3, 64), (2, 2, 3), (3, 1, 2), (4, 5, 64)])
def test_replay_matches_reference_loop(seed, stride, bank_max):
    rng = np.random.default_rng(seed)
    n = 240
    target = unit(rng.normal(size=512))
    sc = make_scenario(rng, n, target)
    cfg = PrescanParams(prescan_stride=stride, prescan_max_width=10 ** 6, prescan_bank_max=bank_max, prescan_fd_add=0.3,
                        prescan_add_cooldown_samples=2, face_quality_min=50.0, prescan_min_segment_sec=0.25,
                        prescan_pad_sec=0.1, prescan_exit_cooldown_sec=0.2, prescan_boundary_refine_sec=0.0)
    ref_feat = unit(target + rng.normal(0, 0.03, 512))[None]

    def frame(i):
        if i >= n:
            return None
        a = np.zeros((1, 1, 3), np.uint8)
        a[0, 0, 0], a[0, 0, 1] = i % 256, i // 256
        return a

    olog = []
    OP.prescan(frame, 24, n, FakeFace(sc), ref_feat, cfg, log=olog)

    records, P, Fl = to_records(sc)
    idxs = PS.sample_indices(n, stride)
    results = {}
    for native in (True, False):       # pcb_replay (libpcb200, flat arrays) and the pure-Python statement of the same loop
        face = FakeFace(sc)
        glog = []
        trk, bank = PS.replay(records, None, (P, Fl), idxs, 24, n, face, ref_feat, cfg, log=glog, distances=NumpyDistances(P, Fl),
                              native=native)
        assert len(glog) == len(olog)
        for g, o in zip(glog, olog):
            assert g["idx"] == o["idx"] and g["skip"] == o["skip"] and g["active_before"] == o["active_before"], (native, g, o)
            assert g["nfaces"] == o["nfaces"] and abs(g["best"] - o["best"]) < 1e-6, (native, g, o)
        results[native] = (trk.finish(), len(bank), face._prescan_rr, face._frame_idx, face._no_face_streak,
                           [(g["idx"], g["skip"]) for g in glog], np.array([g["best"] for g in glog]), bank.array())
    assert results[True][:6] == results[False][:6]
    # the native bank normalises with its own float32 summation order: rows (and with them distances) agree to rounding
    np.testing.assert_allclose(results[True][6], results[False][6], rtol=0, atol=1e-6)
    np.testing.assert_allclose(results[True][7], results[False][7], rtol=0, atol=2e-7)
    assert any(r["active_before"] for r in olog) and any(r["skip"] for r in olog)


def test_bank_update_matches_oracle():
    rng = np.random.default_rng(7)
    cfg = PrescanParams(prescan_bank_max=4)
    bank = PS.RefBank(cfg)
    olist, oarr = [], None
    base = unit(rng.normal(size=512))
    actions = set()
    for t in range(200):
        v = unit(base + rng.normal(0, rng.choice([0.005, 0.05, 0.3]), 512)) * np.float32(rng.uniform(0.5, 2.0))
        q = float(rng.uniform(0, 1200))
        a = bank.offer(v, q)
        oarr, oa, _ = OP.bank_update(olist, oarr, v, q, cfg)
        assert a == oa, (t, a, oa)
        actions.add(a)
        assert np.array_equal(bank.array(), oarr)
    assert {"added", "dup", "replaced", "skip"} <= actions


def test_span_tracker_and_bridge():
    cfg = PrescanParams(prescan_stride=2, prescan_pad_sec=0.25, prescan_min_segment_sec=0.5)
    trk = PS.SpanTracker(cfg, 24, 400)
    seq = [9.0] * 5 + [0.3] * 10 + [0.6] * 3 + [9.0] * 30 + [0.4] * 20
    for k, b in enumerate(seq):
        trk.observe(2 * k, b)
    spans = trk.finish()
    assert spans and all(e - s + 1 >= 12 for s, e in spans)
    assert PS.bridge_spans([(0, 10), (15, 30), (100, 120)], 5) == OP._bridge([(0, 10), (15, 30), (100, 120)], 5) == [(0, 30), (100, 120)]


def test_cache_layout_is_the_references(tmp_path):
    cfg = PrescanParams(video=str(tmp_path / "clip.mp4"), ref=str(tmp_path / "a.png") + ";" + str(tmp_path / "b.png"))
    (tmp_path / "clip.mp4").write_bytes(b"x" * 10)
    (tmp_path / "a.png").write_bytes(b"y")
    m1, m2 = PS.cache_meta(cfg, 23.976024, 1000), OP.cache_meta(cfg, 23.976024, 1000)
    assert m1 == m2 and len(m1["key"]) == 64 and set(m1["settings"]) == set(OP.CACHE_KEYS) and len(OP.CACHE_KEYS) == 34
    spans = [(5, 90), (200, 260)]
    bank = np.random.default_rng(0).normal(size=(3, 512)).astype(np.float32)
    p = PS.save_cache(cfg, 23.976024, 1000, spans, bank, root=tmp_path / "cache")
    assert p.name == m1["key"] + ".npz"
    with np.load(p, allow_pickle=False) as z:
        assert set(z.files) == {"meta", "spans", "ref_face_feat", "has_ref"}
        assert z["spans"].dtype == np.int64 and z["spans"].shape == (2, 2)
        assert z["ref_face_feat"].dtype == np.float32 and z["has_ref"].dtype == np.uint8 and z["has_ref"].shape == (1,)
        assert json.loads(str(z["meta"].item()))["key"] == m1["key"]
    for loader in (PS.load_cache, OP.load_cache):
        hit, s, b, _ = loader(cfg, 23.976024, 1000, tmp_path / "cache")
        assert hit and s == spans and np.array_equal(b, bank)
    q = OP.save_cache(cfg, 23.976024, 1000, spans, None, tmp_path / "c2")
    hit, s, b, _ = PS.load_cache(cfg, 23.976024, 1000, tmp_path / "c2")
    assert hit and s == spans and b is None
    cfg.prescan_cache_mode = "off"
    assert PS.save_cache(cfg, 23.976024, 1000, spans, bank, root=tmp_path / "c3") is None
    assert PS.load_cache(cfg, 23.976024, 1000, tmp_path / "cache")[0] is False


class _FakeEngine:
    """CPU stand-in for Engine.embed / Engine.match: a fixed random projection of the chip (mirrored for the flip variant)."""
    stream = None

    def __init__(self, bank):
        import torch
        g = torch.Generator().manual_seed(7)
        self.proj = torch.randn(8 * 8 * 3, 512, generator=g)
        self.bank = torch.as_tensor(bank, dtype=torch.float32)

    def _e(self, chips, flip):
        import torch
        x = chips.to(torch.float32)
        if flip:
            x = torch.flip(x, dims=[2])
        return x.reshape(x.shape[0], -1) @ self.proj

    def embed(self, chips, f, flip):
        only = flip == "only"
        return (None if only else self._e(chips[:f], False)), (self._e(chips[:f], True) if flip else None)

    def match(self, emb, emb_flip, use_flip, f, want_feat=True):
        import torch
        v = emb[:f] + (emb_flip[:f] if emb_flip is not None else 0)
        v = v / v.norm(dim=1, keepdim=True).clamp_min(1e-6)
        sim = (v @ self.bank.T).max(dim=1).values
        return v, sim, torch.zeros(f, dtype=torch.int32)

    def sync(self):
        pass


def test_lazy_face_table_flips_on_demand(monkeypatch):
    """FaceTable in lazy mode: e(x) up front in runs of EMBED_RUN, e(flip x) only for the rows ensure_flip is asked for."""
    import torch
    rng = np.random.default_rng(3)
    bank = rng.standard_normal((3, 512)).astype(np.float32)
    bank /= np.linalg.norm(bank, axis=1, keepdims=True)
    monkeypatch.setattr(PS.FaceTable, "EMBED_RUN", 16)
    batches = [torch.as_tensor(rng.integers(0, 256, (k, 8, 8, 3), dtype=np.uint8)) for k in (7, 13, 5, 22, 9)]
    eng = _FakeEngine(bank)
    lazy, eager = PS.FaceTable(lazy=True), PS.FaceTable(lazy=False)
    for b in batches:
        lazy.queue(eng, b, b.shape[0])
        eager.queue(eng, b, b.shape[0])
    lazy.finalize(eng)
    eager.finalize(eng)
    assert lazy.count == eager.count == 56 and not lazy.flip_ready.any() and lazy.flip_passes == 0
    want = np.array([1, 5, 6, 0, 8, 14, 15, 17, 19, 20, 30, 2, 33, 40, 41, 5])
    assert lazy.ensure_flip(eng, want) and not lazy.ensure_flip(eng, want[:4])
    done = np.unique(want)
    assert lazy.flip_ready[done].all() and lazy.flip_ready.sum() == len(done) and lazy.flip_passes == len(done)
    np.testing.assert_allclose(lazy.flip_host[done], eager.flip[torch.as_tensor(done)].numpy(), rtol=0, atol=1e-6)
    np.testing.assert_allclose(lazy.plain.numpy(), eager.plain.numpy(), rtol=0, atol=0)


def _native_bank(cfg, rows=None):
    import ctypes as C
    from person_capture_b200 import _lib as L
    lib = L.load()
    bc = PS.bank_cfg_of(cfg)
    r = None if rows is None else np.ascontiguousarray(rows, np.float32)
    nb = lib.pcb_bank_create(C.byref(bc), r.ctypes.data_as(C.c_void_p) if r is not None else None, 0 if r is None else len(r))
    assert nb
    return lib, nb


def test_native_bank_matches_python_bank():
    """pcb_bank_offer (the bank the native replay uses) takes the same decisions as RefBank.offer / the oracle's bank_update
    on a sequence that appends, rejects duplicates, replaces and skips; rows agree to float32 rounding."""
    import ctypes as C
    from person_capture_b200 import _lib as L
    rng = np.random.default_rng(7)
    for cap in (4, 64):
        cfg = PrescanParams(prescan_bank_max=cap)
        bank = PS.RefBank(cfg)
        lib, nb = _native_bank(cfg)
        base = unit(rng.normal(size=512))
        actions = set()
        for t in range(300):
            v = (unit(base + rng.normal(0, rng.choice([0.005, 0.05, 0.3]), 512)) * np.float32(rng.uniform(0.5, 2.0))).astype(np.float32)
            if t == 17:
                v = np.zeros(512, np.float32)
            q = float(rng.uniform(0, 1200))
            a = bank.offer(v, q)
            slot = C.c_int32(-1)
            na = L.BANK_ACTIONS[lib.pcb_bank_offer(nb, v.ctypes.data_as(C.c_void_p), q, C.byref(slot))]
            assert na == a, (cap, t, na, a)
            actions.add(a)
            n = lib.pcb_bank_rows(nb)
            assert n == len(bank) and lib.pcb_bank_version(nb) == bank.version
            if n:
                got = np.ctypeslib.as_array(lib.pcb_bank_data(nb), shape=(n, 512))
                np.testing.assert_allclose(got, bank.array(), rtol=0, atol=2e-7)
                if a in ("added", "replaced"):
                    np.testing.assert_allclose(got[slot.value], v / np.linalg.norm(v), rtol=0, atol=2e-7)
        lib.pcb_bank_destroy(nb)
        assert {"added", "dup", "skip"} <= actions and (cap == 64 or "replaced" in actions)


def test_replay_callback_failure_is_raised():
    """An exception inside a ctypes callback must abort the native replay and surface (ctypes would print and swallow it)."""
    rng = np.random.default_rng(0)
    n = 240
    target = unit(rng.normal(size=512))
    sc = make_scenario(rng, n, target)
    cfg = PrescanParams(prescan_stride=1, prescan_max_width=10 ** 6, prescan_fd_add=0.3, prescan_add_cooldown_samples=2,
                        face_quality_min=50.0, prescan_min_segment_sec=0.25, prescan_pad_sec=0.1)
    ref_feat = unit(target + rng.normal(0, 0.03, 512))[None]
    records, P, Fl = to_records(sc)

    class Exploding(NumpyDistances):
        calls = 0

        def get(self, bank):
            Exploding.calls += 1
            if Exploding.calls >= 3:
                raise RuntimeError("matcher lost its device")
            return super().get(bank)

    with pytest.raises(RuntimeError, match="matcher lost its device"):
        PS.replay(records, None, (P, Fl), PS.sample_indices(n, 1), 24, n, FakeFace(sc), ref_feat, cfg, distances=Exploding(P, Fl))
    assert Exploding.calls == 3


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("carry", [False, True])
def test_vectorised_flip_prediction_equals_the_loop(seed, carry):
    """The flat-record form of the flip prediction (used inside the timed step) selects exactly the rows of the per-sample loop."""
    rng = np.random.default_rng(seed)
    cfg = PrescanParams()
    fps, n = 24, 160
    records, rows_next = {}, 0
    for i in range(n):
        rec = PS.SampleRecord(i)
        k = int(rng.integers(0, 4))
        if k:
            rec.up = PS._Variant(np.zeros((k, 4), np.int32), np.ones(k), np.arange(rows_next, rows_next + k))
            rows_next += k
        else:
            for deg in (90, 270):
                if rng.random() < 0.3:
                    rec.hits[deg] = rec.heavy_raw[deg] = 1
                    rec.heavy[deg] = PS._Variant(np.zeros((1, 4), np.int32), np.ones(1), np.arange(rows_next, rows_next + 1))
                    rows_next += 1
        records[i] = rec
    fd = rng.uniform(0.2, 1.2, rows_next)
    idxs = list(range(n))
    meta, _, _ = PS.encode_records(records, idxs, rows_next)
    want = np.sort(PS._predict_flip_rows(records, idxs, fd, cfg, fps, carry_in=carry))
    got = np.sort(PS._predict_flip_rows_meta(meta, fd, cfg, fps, carry_in=carry))
    assert np.array_equal(got, want) and len(want) > 0


@pytest.mark.parametrize("seed,bank_max", [(0, 64), (2, 3), (11, 5)])
def test_replay_dup_filter_is_a_pure_shortcut(seed, bank_max, monkeypatch):
    """pcb_replay skips bank offers whose live distance says "certain duplicate" (similarity >= dedup + 1e-4).  With the
    shortcut switched off (PCB_REPLAY_DUP_FILTER=0) every offer is evaluated: spans, bank and log must not change."""
    rng = np.random.default_rng(seed)
    n = 240
    target = unit(rng.normal(size=512))
    sc = make_scenario(rng, n, target)
    cfg = PrescanParams(prescan_stride=1, prescan_max_width=10 ** 6, prescan_bank_max=bank_max, prescan_fd_add=0.3,
                        prescan_add_cooldown_samples=0, face_quality_min=50.0, prescan_min_segment_sec=0.25,
                        prescan_pad_sec=0.1, prescan_exit_cooldown_sec=0.2, prescan_boundary_refine_sec=0.0)
    ref_feat = unit(target + rng.normal(0, 0.03, 512))[None]
    records, P, Fl = to_records(sc)
    idxs = PS.sample_indices(n, 1)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("PCB_REPLAY_DUP_FILTER", mode)
        face, glog = FakeFace(sc), []
        trk, bank = PS.replay(records, None, (P, Fl), idxs, 24, n, face, ref_feat, cfg, log=glog, distances=NumpyDistances(P, Fl), native=True)
        res[mode] = (trk.finish(), bank.version, [(g["idx"], g["skip"], g["active_before"], g["nfaces"], g["best"]) for g in glog], bank.array())
    assert res["0"][:3] == res["1"][:3]
    assert np.array_equal(res["0"][3], res["1"][3])
    assert res["0"][1] >= 3


def test_bank_identity_survives_address_reuse():
    """Engine.set_bank skips the upload when the bank's (serial, version) is already on the device.  The identity must not be
    id(bank): CPython reuses the address of a freed object, and a new bank then passed for the previous one."""
    cfg = PrescanParams()
    seen = set()
    for _ in range(200):
        b = PS.RefBank(cfg, unit(np.random.default_rng(0).normal(size=512))[None])
        assert (b.serial, b.version) not in seen
        seen.add((b.serial, b.version))
        del b


@pytest.mark.parametrize("seed,trim,skip_trailing", [(0, True, True), (3, True, False), (5, False, False), (12, True, False)])
def test_refine_from_the_face_table_equals_probing_every_frame(seed, trim, skip_trailing):
    """_refine_edges_batched with `known` (stride 1: every probe frame is a sample of the main scan) reads the probe results
    from the face table -- variant choice in "full" mode (upright, else 90 then 270 where the probe hit and the heavy pass found
    a face), flip-TTA distance to the final bank -- and must return what the oracle's _refine_edges returns when it extracts
    every probe frame again."""
    import types
    rng = np.random.default_rng(seed)
    n = 200
    target = unit(rng.normal(size=512))
    sc = make_scenario(rng, n, target)
    cfg = PrescanParams(prescan_stride=1, prescan_max_width=10 ** 6, prescan_min_segment_sec=0.25, prescan_pad_sec=0.25,
                        prescan_boundary_refine_sec=0.5, prescan_trim_pad=trim, prescan_skip_trailing_refine=skip_trailing,
                        prescan_fd_enter=0.45)
    bank_rows = np.stack([unit(target + rng.normal(0, 0.03, 512)), unit(target + rng.normal(0, 0.05, 512))])
    spans = [(10, 45), (70, 120), (150, n - 1)]

    def frame(i):
        if i < 0 or i >= n:
            return None
        a = np.zeros((1, 1, 3), np.uint8)
        a[0, 0, 0], a[0, 0, 1] = i % 256, i // 256
        return a

    want = OP._refine_edges(list(spans), frame, 24, n, FakeFace(sc), bank_rows, None, cfg, int(round(cfg.prescan_min_segment_sec * 24)),
                            10 ** 6, float(cfg.prescan_fd_enter))

    records, P, Fl = to_records(sc)
    idxs = PS.sample_indices(n, 1)
    meta, _, _ = PS.encode_records(records, idxs, len(P))
    import torch
    table = types.SimpleNamespace(plain=torch.as_tensor(P), flip=torch.as_tensor(Fl), count=len(P), lazy=False)
    bank = PS.RefBank(cfg, bank_rows)
    dist_calls = []

    class Dist:
        def get(self, b):
            dist_calls.append(b.version)
            return NumpyDistances(P, Fl).get(b)

        def invalidate(self):
            pass

    known = dict(pos={int(j): k for k, j in enumerate(idxs)}, meta=meta, table=table, dist=Dist())
    clip = types.SimpleNamespace(total_frames=n)
    trk = PS.SpanTracker(cfg, 24, n)
    got = PS._refine_edges_batched(list(spans), clip, 24, FakeFace(sc), bank, None, cfg, trk, 16, known=known)
    assert got == want
    assert dist_calls                                  # the table was consulted; compute_superset (GPU) was never needed
    assert want != spans or not trim                   # the case moves at least one edge when trimming is on


def test_video_file_clip_decodes_what_cv2_decodes(tmp_path):
    """VideoFileClip (CPU decode through OpenCV's FFmpeg reader, the reference's default SDR reader): sequential reads, reads
    into a caller's buffer and random access all return the frames a plain cv2.VideoCapture walk returns."""
    import cv2
    path = str(tmp_path / "clip.avi")
    rng = np.random.default_rng(5)
    w, h, n = 96, 64, 24
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 24.0, (w, h))
    if not vw.isOpened():
        pytest.skip("no MJPG writer in this OpenCV build")
    for i in range(n):
        f = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 2.0)
        cv2.putText(f, str(i), (5, 40), cv2.FONT_HERSHEY_SIMPLEX, 1.0, (255, 255, 255), 2)
        vw.write(f)
    vw.release()
    cap = cv2.VideoCapture(path)
    want = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        want.append(f)
    cap.release()
    assert len(want) == n
    clip = PS.VideoFileClip(path)
    assert clip.total_frames == n and (clip.height, clip.width) == (h, w) and clip.host_resident and abs(clip.fps - 24.0) < 1e-6
    for i in range(n):
        assert np.array_equal(clip.host(i), want[i]), i
    assert clip.seeks == 0 and clip.decoded == n
    buf = np.zeros((h, w, 3), np.uint8)
    for i in (17, 3, 4, 23, 0):                       # random access (MJPG is intra-only: seeks are exact), decode into a buffer
        assert clip._read_into(i, buf) is buf and np.array_equal(buf, want[i]), i
    assert clip.seeks == 4                             # 3 -> 4 was sequential
    with pytest.raises(IndexError):
        clip.host(n)
    clip.close()
    with pytest.raises(PS.L.PcbError):
        PS.VideoFileClip(str(tmp_path / "missing.mp4"))
