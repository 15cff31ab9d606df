"""CPU: the oracle (and the product's host logic) against vectors produced by the UNMODIFIED reference.

tests/golden/reference_golden.npz was written by tests/golden/make_reference_golden.py, which imports
`/root/reference/person_capture` in the build container (tests/golden/ref_harness.py) and runs the reference's own
`FaceEmbedder` / `Processor` methods on the seeded inputs of tests/golden/ref_golden_script.py.  Nothing here needs the
reference at run time.  Bars: bit-exact for bytes / integers / decisions; float64 quantities to 1e-12; float32 feature
vectors to 1e-6 (the stand-in ArcFace map runs through BLAS).
"""
import json
import os
import sys
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import pcb_test_helpers as H  # noqa: E402
import ref_golden_script as S  # noqa: E402
from oracle import face_embedder as OF  # noqa: E402
from oracle import prescan as OP  # noqa: E402
from person_capture_b200 import prescan as PS  # noqa: E402
from person_capture_b200.params import PrescanParams  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "reference_golden.npz"), allow_pickle=False)


def crc(a) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------- units
def test_units_match_reference():
    cases = S.unit_cases()
    n_none = 0
    for k, (crop, pts) in enumerate(cases):
        c = OF.canon_5pts(pts.copy())
        assert (c is not None) == bool(G["unit_canon_ok"][k]), k
        if c is not None:
            assert np.array_equal(np.asarray(c, np.float32), G["unit_canon"][k]), k
        else:
            n_none += 1
        chip_a = OF.align_by_5pts(crop, c if c is not None else pts)
        chip_r = OF.upright_by_eye_roll(crop, pts)
        assert np.array_equal(chip_a, G["unit_align"][k]), ("align", k)
        assert np.array_equal(chip_r, G["unit_roll"][k]), ("roll", k)
        q = G["unit_quality"][k]
        assert abs(OF.face_quality(chip_a) - q[0]) <= 1e-12 * max(1.0, q[0]) and abs(OF.face_quality(chip_r) - q[1]) <= 1e-12 * max(1.0, q[1])
        assert crc(OF.arcface_preprocess(chip_a)) == int(G["unit_pre_crc"][k]), ("preprocess", k)
    assert 4 <= n_none <= len(cases) - 4          # both canon outcomes are exercised
    import cv2
    assert np.array_equal(OF.arcface_preprocess(cases[0][0][:40, :36]), G["unit_pre_small"][0])
    assert np.array_equal(OF.arcface_preprocess(cv2.resize(cases[1][0], (150, 170), interpolation=cv2.INTER_LINEAR)), G["unit_pre_large"][0])


# ------------------------------------------------------------------------------------------------- extract
class ReplayScrfd:
    """Returns the detections the reference's run recorded, and demands that the caller made the SAME detector call:
    same input image (CRC + shape), same input size, same threshold, in the same order."""

    def __init__(self, prefix="ex"):
        self.prefix = prefix
        self.meta, self.thresh = G[prefix + "_call_meta"], G[prefix + "_call_thresh"]
        self.off = np.concatenate([[0], np.cumsum(self.meta[:, 5])]).astype(np.int64)
        self.k = 0
        self.cur_extract = 0
        self.det_thresh = 0.5

    def detect(self, img, input_size=None, **kw):
        k = self.k
        assert k < len(self.meta), "more SCRFD passes than the reference made"
        ex, c, h, w, size, n = [int(v) for v in self.meta[k]]
        assert ex == self.cur_extract, f"pass {k}: the reference made this pass in extract call {ex}, not {self.cur_extract}"
        assert (img.shape[0], img.shape[1], int(input_size[0])) == (h, w, size), (k, img.shape, input_size, (h, w, size))
        assert abs(float(self.det_thresh) - float(self.thresh[k])) < 1e-12, (k, self.det_thresh, self.thresh[k])
        assert crc(img) == c, f"pass {k}: detector input image differs from the reference's"
        self.k += 1
        a, b = self.off[k], self.off[k + 1]
        return G[self.prefix + "_det"][a:b].copy(), G[self.prefix + "_kps"][a:b].copy()


def test_extract_matches_reference():
    rep = ReplayScrfd()
    O = OF.FaceEmbedderOracle(rep, H.ProjArcface(), conf=0.5)
    counts, state = G["ex_out_count"], G["ex_state"]
    face_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    n_ex = 0
    chip_k = 0
    for op in S.extract_script():
        if op[0] == "knob":
            getattr(O, op[1])(**op[2])
            continue
        if op[0] == "attr":
            setattr(O, op[1], op[2])
            continue
        rep.cur_extract = n_ex
        frame = S.frame_of(op[1])
        faces = O.extract(frame) if op[2] is None else O.extract(frame, imgsz=op[2])
        # every pass the reference made in this call has been consumed
        want_passes = int((G["ex_call_meta"][:, 0] <= n_ex).sum())
        assert rep.k == want_passes, (n_ex, op, rep.k, want_passes)
        a, b = face_off[n_ex], face_off[n_ex + 1]
        assert len(faces) == b - a, (n_ex, op, len(faces), b - a)
        for j, f in enumerate(faces):
            assert np.array_equal(np.asarray(f["bbox"], np.int32), G["ex_bbox"][a + j]), (n_ex, j)
            assert abs(float(f["quality"]) - G["ex_quality"][a + j]) <= 1e-12 * max(1.0, G["ex_quality"][a + j]), (n_ex, j)
            np.testing.assert_allclose(np.asarray(f["feat"], np.float32), G["ex_feat"][a + j], rtol=0, atol=1e-6)
        # chips in pre-sort order, as handed to ArcFace
        chips = O.last_chips if len(faces) else []
        assert len(chips) == b - a
        for c in chips:
            assert crc(c) == int(G["ex_chip_crc"][chip_k]), (n_ex, chip_k)
            if chip_k < len(G["ex_chips"]):
                assert np.array_equal(c, G["ex_chips"][chip_k])
            chip_k += 1
        st = state[n_ex]
        assert (O._no_face_streak, O._rot_cycle, O._prescan_rr, O._frame_idx, O._last_face_idx) == tuple(int(v) for v in st[:5]), (n_ex, op)
        n_ex += 1
    assert rep.k == len(G["ex_call_meta"]) and n_ex == len(counts)
    assert counts.sum() >= 30 and (counts == 0).sum() >= 10       # both outcomes well represented
    sizes = set(int(v) for v in G["ex_call_meta"][:, 4])
    assert len(sizes) >= 5, sizes                                    # upright, shrunk, probe, heavy-90/180, explicit imgsz


# ------------------------------------------------------------------------------------------------- bank
def _cfg(over):
    return PrescanParams(**over)


@pytest.mark.parametrize("name", sorted(S.bank_cases().keys()))
def test_bank_update_matches_reference(name):
    over, feats, quals, seed_rows = S.bank_cases()[name]
    cfg = _cfg(over)
    want_a, want_i, want_fd, want_final = G[f"bank_{name}_actions"], G[f"bank_{name}_idx"], G[f"bank_{name}_fd"], G[f"bank_{name}_final"]
    # oracle
    olist = [r.copy() for r in seed_rows]
    oarr = np.vstack(olist).astype(np.float32) if olist else None
    # product (Python RefBank) and native pcb_bank
    pbank = PS.RefBank(cfg, np.vstack(seed_rows) if seed_rows else None)
    import ctypes as C
    from person_capture_b200 import _lib as L
    lib = L.load()
    bc = PS.bank_cfg_of(cfg)
    seed = np.ascontiguousarray(np.vstack(seed_rows), np.float32) if seed_rows else None
    nb = lib.pcb_bank_create(C.byref(bc), seed.ctypes.data_as(C.c_void_p) if seed is not None else None, 0 if seed is None else len(seed))
    assert nb
    try:
        for t, (v, q) in enumerate(zip(feats, quals)):
            fd = OP.fd_min(v, oarr)
            assert abs(fd - want_fd[t]) <= 1e-12, (t, fd, want_fd[t])
            oarr, act, idx = OP.bank_update(olist, oarr, v, float(q), cfg)
            assert S.ACTIONS.index(act) == int(want_a[t]) and (-1 if idx is None else int(idx)) == int(want_i[t]), (t, act, idx)
            assert pbank.offer(v, float(q)) == act, t
            slot = C.c_int32(-1)
            vv = np.ascontiguousarray(v, np.float32)
            code = lib.pcb_bank_offer(nb, vv.ctypes.data_as(C.c_void_p), float(q), C.byref(slot))
            assert L.BANK_ACTIONS[code] == act, (t, code, act)
            if act == "replaced":
                assert slot.value == int(want_i[t])
        assert np.array_equal(oarr, want_final)
        assert np.array_equal(pbank.array(), want_final)
        rows = lib.pcb_bank_rows(nb)
        nat = np.ctypeslib.as_array(lib.pcb_bank_data(nb), shape=(rows, 512)).copy()
        np.testing.assert_allclose(nat, want_final, rtol=0, atol=2e-7)       # native rows: own float32 summation order in the norm
    finally:
        lib.pcb_bank_destroy(nb)
    assert set(int(a) for a in want_a) == {0, 1, 2, 3} or name != "cap4"


def test_fd_min_corner_cases_match_reference():
    v = S.unit_vec(7)
    got = [OP.fd_min(None, np.eye(4, 512, dtype=np.float32)), OP.fd_min(v, None), OP.fd_min(v, np.zeros((0, 512), np.float32)),
           OP.fd_min(v, v), OP.fd_min(3.0 * v, np.stack([S.unit_vec(8), v]))]
    np.testing.assert_allclose(got, G["fdmin_corner"], rtol=0, atol=1e-12)


# ------------------------------------------------------------------------------------------------- prescan
@pytest.mark.parametrize("name", sorted(S.prescan_cases().keys()))
def test_prescan_matches_reference(name):
    case = S.prescan_cases()[name]
    cfg = _cfg(case["cfg"])
    n, fps = case["n"], case["fps"]
    sc, ref = S.prescan_inputs(case)
    cap = S.PipeLikeCap(n, case.get("h", 2), case.get("w", 2))
    face = S.RecordingFakeFace(sc)
    log = []
    spans, bank = OP.prescan(cap.frame, fps, n, face, ref, cfg, log=log)
    want_spans = [tuple(int(v) for v in r) for r in G[f"ps_{name}_spans"]]
    assert [tuple(s) for s in spans] == want_spans
    want_bank = G[f"ps_{name}_bank"]
    got_bank = np.zeros((0, 512), np.float32) if bank is None else np.asarray(bank, np.float32).reshape(-1, 512)
    assert np.array_equal(got_bank, want_bank)
    # the face object saw the same calls in the same order with the same knob state and frame width (main loop + refine)
    assert np.array_equal(np.array(face.calls, np.int64).reshape(-1, 4), G[f"ps_{name}_calls"])
    st = G[f"ps_{name}_final_state"]
    assert (face.conf, float(face.rot_adaptive), float(face._prescan_escalate), float(face._prescan_rr_mode == "full"),
            float(face._prescan_rr), float(face._frame_idx)) == tuple(float(v) for v in st[:6])

    # product: native pcb_replay and its Python statement reproduce the main loop of the reference (the refine stage needs the
    # GPU engine and is checked against the oracle in tests/test_gpu_e2e.py)
    if case.get("w", 2) > int(case["cfg"].get("prescan_max_width", 10 ** 6)):
        return                                    # frame resize happens in the GPU stage of the product
    from scenario_helpers import to_records
    from test_cpu_host_logic import NumpyDistances
    records, P, Fl = to_records(sc)
    idxs = PS.sample_indices(n, max(1, int(cfg.prescan_stride)))
    n_main = sum(1 for r in log if not r["skip"])
    for native in (True, False):
        f2 = S.RecordingFakeFace(sc)
        glog = []
        trk, pbank = PS.replay(records, None, (P, Fl), idxs, fps, n, f2, ref, cfg, log=glog, distances=NumpyDistances(P, Fl), native=native)
        assert [(g["idx"], g["skip"], g["active_before"]) for g in glog] == [(o["idx"], o["skip"], o["active_before"]) for o in log]
        main_calls = G[f"ps_{name}_calls"][:n_main]
        assert [g["idx"] for g in glog if not g["skip"]] == [int(v) for v in main_calls[:, 0]]
        assert [int(g["active_before"]) for g in glog if not g["skip"]] == [int(v) for v in main_calls[:, 2]]
        pb = pbank.array()
        pb = np.zeros((0, 512), np.float32) if pb is None else pb
        np.testing.assert_allclose(pb, want_bank if len(want_bank) else pb, rtol=0, atol=2e-7)
        assert len(pb) == len(want_bank) or (ref is not None and len(want_bank) == len(np.atleast_2d(ref)) and len(pb) == len(want_bank))


def test_prescan_with_real_embedder_matches_reference():
    """Processor._prescan driving FaceEmbedder.extract (both the reference's own code) over a rendered 640x360 clip, against
    the oracle's prescan driving the oracle's embedder: same detector calls in the same order (frame downscale, probe sizes,
    rotation cadence under rr / full mode, thresholds), same spans after bridge + refine, same grown bank, same final state."""
    rep = ReplayScrfd("pf")
    O = OF.FaceEmbedderOracle(rep, H.ProjArcface(), conf=0.5)
    frames, ref_img = S.full_clip_frames()
    cfg = _cfg(S.FULL_CFG)
    rfaces = O.extract(ref_img)
    ref = np.asarray(max(rfaces, key=lambda f: f["quality"])["feat"], np.float32)[None]
    np.testing.assert_allclose(ref, G["pf_ref"], rtol=0, atol=1e-6)
    spans, bank = OP.prescan(lambda i: frames[i] if 0 <= i < len(frames) else None, S.FULL_FPS, S.FULL_N, O, G["pf_ref"], cfg)
    assert rep.k == len(G["pf_call_meta"])                       # every pass of the reference's run, none more
    assert [tuple(int(v) for v in sp) for sp in spans] == [tuple(int(v) for v in r) for r in G["pf_spans"]] and len(spans) >= 2
    np.testing.assert_allclose(np.asarray(bank, np.float32), G["pf_bank"], rtol=0, atol=1e-6)
    assert len(G["pf_bank"]) > 1
    st = G["pf_state"]
    assert (O._no_face_streak, O._rot_cycle, O._prescan_rr, O._frame_idx, O._last_face_idx, int(O._fast_prescan),
            int(O._prescan_escalate), int(O.rot_adaptive)) == tuple(int(v) for v in st)


# ------------------------------------------------------------------------------------------------- cache
def test_cache_key_and_file_match_reference(tmp_path):
    video, ref = S.cache_files()
    st = os.stat(video)
    if st.st_mtime_ns != 1_700_000_000_123_456_000:
        pytest.skip("file system does not keep the mtime the cache key fingerprints")
    try:
        for mod in (OP, PS):
            cfg = _cfg(dict(S.CACHE_CFG, video=video, ref=ref))
            meta = mod.cache_meta(cfg, 23.976, 4321)
            key_json = json.dumps({k: v for k, v in meta.items() if k != "key"}, sort_keys=True, separators=(",", ":"))
            assert key_json == str(G["cache_key_json"]), mod.__name__
            assert meta["key"] == str(G["cache_key"])
            cfg2 = _cfg(dict(S.CACHE_CFG, video=video, ref="", prescan_stride=7))
            assert mod.cache_meta(cfg2, 30.0, 100)["key"] == str(G["cache2_key"])
        # the reference's own file, byte for byte, through both loaders
        root = tmp_path / "cache"
        root.mkdir()
        (root / str(G["cache_path_name"])).write_bytes(G["cache_file_bytes"].tobytes())
        cfg = _cfg(dict(S.CACHE_CFG, video=video, ref=ref))
        for loader in (OP.load_cache, PS.load_cache):
            hit, spans, bank, _ = loader(cfg, 23.976, 4321, root)
            assert hit and spans == [(10, 200), (400, 4320)]
            assert np.array_equal(bank, G["cache_file_ref"])
        # and our writer produces the same arrays / dtypes / meta string
        p = PS.save_cache(cfg, 23.976, 4321, [(10, 200), (400, 4320)], G["cache_file_ref"], root=tmp_path / "c2")
        assert p.name == str(G["cache_path_name"])
        with np.load(p, allow_pickle=False) as z:
            assert sorted(z.files) == [str(k) for k in G["cache_file_keys"]]
            assert [str(z[k].dtype) for k in sorted(z.files)] == [str(d) for d in G["cache_file_dtypes"]]
            assert str(z["meta"].item()) == str(G["cache_file_meta"])
            assert np.array_equal(z["spans"], G["cache_file_spans"]) and np.array_equal(z["has_ref"], G["cache_file_has_ref"])
        p2 = PS.save_cache(_cfg(dict(S.CACHE_CFG, video=video, ref="", prescan_stride=7)), 30.0, 100, [], None, root=tmp_path / "c3")
        with np.load(p2, allow_pickle=False) as z:
            assert tuple(z["ref_face_feat"].shape) == tuple(int(v) for v in G["cache2_ref_shape"])
            assert tuple(z["spans"].shape) == tuple(int(v) for v in G["cache2_spans_shape"])
            assert np.array_equal(z["has_ref"], G["cache2_has_ref"])
    finally:
        import shutil
        shutil.rmtree(S.CACHE_DIR, ignore_errors=True)


# ------------------------------------------------------------------------------------------------- curator
def test_curator_identity_matches_reference():
    from oracle.curator import CuratorIdentityOracle, letterbox_square
    rows, canv = [], []
    for square in (True, False):
        for passed in (False, True):
            face = S.CannedFace()
            C = CuratorIdentityOracle(face, None, det_square=square, id_already_passed=passed)
            C.ref_feat = S.unit_vec(3) * np.float32(1.3)
            for img in S.curator_images():
                best = C.detect_best_face(img)
                if best is None:
                    rows.append([0, 0, 0, 0, 0, 0.0, C.fd_min(None)])
                else:
                    rows.append([1, *[int(v) for v in best["bbox"]], float(best["quality"]), C.fd_min(best["feat"])])
            canv.extend([c for (_, _, c) in face.seen])
    np.testing.assert_allclose(np.array(rows, np.float64), G["cur_rows"], rtol=0, atol=1e-12)
    assert np.array_equal(np.array(canv, np.int64), G["cur_canvas_crc"])          # the detector saw the same canvases
    lb = [letterbox_square(img, 640) for img in S.curator_images()]
    np.testing.assert_allclose(np.array([[sc, dx, dy] for (_, sc, dx, dy) in lb], np.float64), G["cur_lb_meta"], rtol=0, atol=0)
    assert [crc(cv) for (cv, _, _, _) in lb] == [int(v) for v in G["cur_lb_crc"]]
    C.ref_feat = None
    assert C.fd_min(S.unit_vec(1)) == float(G["cur_fd_noref"][0])


# ------------------------------------------------------------------------------------------------- main-pass geometry
def test_main_pass_geometry_matches_reference():
    """Processor._expand_xyxy (lock-face ROI geometry, gui_app.py:4186-4199) and Processor._iou (:3484-3493), oracle and product."""
    from oracle import mainpass as OM
    from person_capture_b200 import mainpass as MP
    cases, pairs = S.geometry_cases()
    for mod in (OM, MP):
        got = np.array([mod.expand_xyxy(b, px, py, W, Hh) for b, px, py, W, Hh in cases], np.int64)
        assert np.array_equal(got, G["geo_expand"]), mod.__name__
    np.testing.assert_allclose([OM.iou_xyxy(a, b) for a, b in pairs], G["geo_iou"], rtol=0, atol=0)
    np.testing.assert_allclose([MP.box_iou(a, b) for a, b in pairs], G["geo_iou"], rtol=0, atol=1e-15)


# ------------------------------------------------------------------------------------------------- reference x ONNX graphs
def test_oracle_matches_reference_running_the_onnx_graphs():
    """`ro_*`: the unmodified reference FaceEmbedder with both sessions executing the EXPORTED ONNX graphs through cv2.dnn
    (make_reference_golden.gen_reference_onnx).  The oracle (torch-CPU executors of the same weights) must see the same faces:
    identical boxes, counters and rotation state; quality and features to the rounding two inference engines differ by."""
    O = H.oracle_embedder(S.RO_SCRFD, S.RO_ARC, conf=0.5)
    O.configure_rotation_strategy(adaptive=False)
    O.set_prescan_fast(True, mode="rr")
    O._prescan_probe_imgsz = 512
    off = np.concatenate([[0], np.cumsum(G["ro_counts"])]).astype(int)
    for k, key in enumerate(S.RO_FRAMES):
        O.set_prescan_hint(escalate=bool(k % 2))
        faces = O.extract(S.ro_frame(key))
        a, b = off[k], off[k + 1]
        assert len(faces) == b - a, k
        for j, f in enumerate(faces):
            assert np.array_equal(np.asarray(f["bbox"], np.int32), G["ro_bbox"][a + j]), (k, j)
            assert abs(f["quality"] - G["ro_quality"][a + j]) <= 1e-3 * max(1.0, G["ro_quality"][a + j]), (k, j)
            assert H.cos(f["feat"], G["ro_feat"][a + j]) >= 0.9999, (k, j)
    assert (O._prescan_rr, O._no_face_streak, O._frame_idx) == tuple(int(v) for v in G["ro_state"])
    assert G["ro_counts"].sum() >= 8 and (G["ro_counts"] == 0).any()


def test_oracle_prescan_matches_reference_prescan_on_the_onnx_graphs():
    """`rp_*`: Processor._prescan x FaceEmbedder (unmodified) x the exported ONNX graphs through cv2.dnn over the 144-frame clip of
    tests/test_gpu_e2e.py::test_prescan_spans_match_oracle.  The oracle (torch executors) builds the same reference bank and keeps
    the same spans and bank."""
    frames, ref_img = S.rp_clip_frames()
    cfg = _cfg(S.RP_CFG)
    ora = H.oracle_embedder(S.RO_SCRFD, S.RO_ARC, conf=cfg.face_det_conf)
    bank0 = OP.build_reference_bank(ora, [ref_img], cfg)
    assert bank0.shape == G["rp_ref"].shape
    for a, b in zip(bank0, G["rp_ref"]):
        assert H.cos(a, b) >= 0.9999
    n_ext = [0]
    inner = ora.extract

    def counted(img, **kw):
        n_ext[0] += 1
        return inner(img, **kw)

    ora.extract = counted
    spans, bank = OP.prescan(lambda i: frames[i] if 0 <= i < len(frames) else None, S.RP_FPS, S.RP_N, ora, bank0, cfg)
    assert [tuple(int(v) for v in sp) for sp in spans] == [tuple(int(v) for v in r) for r in G["rp_spans"]] and len(spans) >= 2
    assert n_ext[0] == int(G["rp_extracts"][0])
    assert np.asarray(bank).shape == G["rp_bank"].shape and len(G["rp_bank"]) > len(G["rp_ref"])
    for a, b in zip(np.asarray(bank), G["rp_bank"]):
        assert H.cos(a, b) >= 0.9999


def test_reference_onnx_vectors_of_the_headline_models_are_present():
    """`rh_*` (SCRFD-10G + iResNet-100 on 960x540 frames, the bench configuration) are consumed by the GPU test
    tests/test_gpu_headline.py::test_extract_960x540_r100_matches_the_reference_running_the_onnx_graphs; here: shape sanity."""
    assert G["rh_counts"].sum() == len(G["rh_bbox"]) == len(G["rh_quality"]) == len(G["rh_feat"]) >= 10
    assert np.allclose(np.linalg.norm(G["rh_feat"], axis=1), 1.0, atol=1e-5)
    assert len(G["rh_counts"]) == len(S.RH_FRAME_IDS)


def test_reference_onnx_prescan_vectors_of_the_headline_models_are_present():
    """`rq_*` (the unmodified _prescan x FaceEmbedder x ONNX graphs with SCRFD-10G + iResNet-100 over a 1080p clip) are consumed
    by tests/test_gpu_headline.py::test_prescan_1080p_r100_matches_the_reference_prescan_on_the_onnx_graphs; here: shape sanity."""
    assert G["rq_spans"].shape[1] == 2 and len(G["rq_spans"]) >= 1 and (G["rq_spans"][:, 1] >= G["rq_spans"][:, 0]).all()
    assert len(G["rq_bank"]) > len(G["rq_ref"]) >= 1
    assert np.allclose(np.linalg.norm(G["rq_bank"], axis=1), 1.0, atol=1e-5)
