"""Generates tests/golden/reference_golden.npz: outputs of the UNMODIFIED reference (`/root/reference/person_capture`),
executed in the build container through tests/golden/ref_harness.py, on seeded inputs.  These vectors pin the oracle
(`oracle/`) -- and, through CPU tests, the host logic of the product -- against the reference itself; they are the answer to
"parity unpinned" for everything on the path that is not a neural-network session:

  units    FaceEmbedder._canon_5pts / _align_by_5pts / _upright_by_eye_roll / _face_quality / _arcface_preprocess
  extract  FaceEmbedder.extract (= _extract_with_scrfd_raw + _arcface_encode): size / rotation policy, TTA, pad probes,
           accumulate + min-size, cross-pass NMS, alignment, flip + sum + normalise, ordering -- driven through a
           script of knob changes and frames; every SCRFD call the reference makes is recorded (input image CRC, size,
           threshold, returned detections) so a test can replay the detector and demand identical calls
  bank     Processor._fd_min, Processor._stream_ref_bank_update over offer sequences (added / dup / replaced / skip)
  prescan  Processor._prescan (state machine, fd9 gate, bank growth, pad / merge, bridge, _refine_edges) over scripted
           per-frame faces; spans, final bank and the knob state handed to the face object at every extract call
  cache    Processor._prescan_cache_meta / _prescan_cache_path / _save_prescan_cache: key, meta JSON, file arrays

The two sessions the reference would open through ONNX Runtime / InsightFace are substituted (ref_harness): SCRFD by the
oracle's fp32 detector on the repo's weights (its answers are RECORDED, so replaying the golden does not depend on that
network's numerics), ArcFace by `pcb_test_helpers.proj_arcface` (a fixed linear map of the pooled blob: deterministic float64).

usage (build container only; /root/reference must exist):  python tests/golden/make_reference_golden.py
"""
import json
import os
import queue
import sys
import tempfile
import zlib

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import pcb_test_helpers as H  # noqa: E402
import ref_golden_script as S  # noqa: E402
import ref_harness as RH  # noqa: E402
from oracle.scrfd_detect import SCRFDOracle  # noqa: E402

OUT = os.path.join(HERE, "reference_golden.npz")


def crc(a) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


class RecordingScrfd:
    """What `self.scrfd` is to the reference: `.det_thresh`, `.detect(img, input_size=)`; records every call."""

    def __init__(self, inner):
        self.inner = inner
        self.det_thresh = 0.5
        self.calls = []

    def detect(self, img, input_size=None, **kw):
        self.inner.det_thresh = float(self.det_thresh)
        det, kps = self.inner.detect(img, input_size=input_size)
        det = np.asarray(det, np.float32).reshape(-1, 5)
        kps = np.zeros((0, 5, 2), np.float32) if kps is None else np.asarray(kps, np.float32).reshape(-1, 5, 2)
        self.calls.append(dict(crc=crc(img), h=img.shape[0], w=img.shape[1], size=int(input_size[0]),
                               thresh=float(self.det_thresh), det=det.copy(), kps=kps.copy()))
        return det, kps


def gen_units(fe, out):
    R = RH.make_reference_embedder(fe, None, H.proj_arcface)
    cases = S.unit_cases()
    canon_ok, canon_out, align_out, roll_out, qual, pre_crc = [], [], [], [], [], []
    for crop, pts in cases:
        c = fe.FaceEmbedder._canon_5pts(pts.copy())
        canon_ok.append(c is not None)
        canon_out.append(np.zeros((5, 2), np.float32) if c is None else c.astype(np.float32))
        chip_a = R._align_by_5pts(crop, c if c is not None else pts)
        chip_r = R._upright_by_eye_roll(crop, pts)
        align_out.append(chip_a)
        roll_out.append(chip_r)
        qual.append([R._face_quality(chip_a), R._face_quality(chip_r)])
        pre_crc.append(crc(R._arcface_preprocess(chip_a)))
    # _arcface_preprocess on non-112 inputs (resize branch, both interpolations)
    pre_small = R._arcface_preprocess(cases[0][0][:40, :36])
    pre_large = R._arcface_preprocess(cv2.resize(cases[1][0], (150, 170), interpolation=cv2.INTER_LINEAR))
    out["unit_canon_ok"] = np.array(canon_ok, np.uint8)
    out["unit_canon"] = np.stack(canon_out)
    out["unit_align"] = np.stack(align_out)
    out["unit_roll"] = np.stack(roll_out)
    out["unit_quality"] = np.array(qual, np.float64)
    out["unit_pre_crc"] = np.array(pre_crc, np.int64)
    out["unit_pre_small"] = pre_small.astype(np.float32)
    out["unit_pre_large"] = pre_large.astype(np.float32)


def gen_extract(fe, out):
    rec = RecordingScrfd(SCRFDOracle(H.oracle_scrfd(S.EXTRACT_SCRFD)))
    R = RH.make_reference_embedder(fe, rec, H.proj_arcface, conf=0.5)
    call_meta, call_thresh, dets, kpss = [], [], [], []
    out_count, bbox, quality, feat, chip_crc, chips, state = [], [], [], [], [], [], []
    n_ex = 0
    for op in S.extract_script():
        if op[0] == "knob":
            getattr(R, op[1])(**op[2])
            continue
        if op[0] == "attr":
            setattr(R, op[1], op[2])
            continue
        assert op[0] == "extract"
        frame = S.frame_of(op[1])
        rec.calls.clear()
        R.arc_sess.blobs.clear()
        faces = R.extract(frame) if op[2] is None else R.extract(frame, imgsz=op[2])
        for c in rec.calls:
            call_meta.append([n_ex, c["crc"], c["h"], c["w"], c["size"], len(c["det"])])
            call_thresh.append(c["thresh"])
            dets.append(c["det"])
            kpss.append(c["kps"])
        out_count.append(len(faces))
        # chips as the reference handed them to ArcFace: blob = (RGB - 127.5) / 127.5, first m rows are the unflipped ones
        m = len(faces)
        blob = R.arc_sess.blobs[0][:m] if (m and R.arc_sess.blobs) else np.zeros((0, 3, 112, 112), np.float32)
        n_blob = R.arc_sess.blobs[0].shape[0] if R.arc_sess.blobs else 0
        ch = np.rint(blob * 127.5 + 127.5).astype(np.uint8).transpose(0, 2, 3, 1)[..., ::-1]
        # `faces` is sorted by (quality, area); blobs are in pre-sort order -> store the sorted outputs and the pre-sort chips
        for f in faces:
            bbox.append(np.asarray(f["bbox"], np.int32))
            quality.append(float(f["quality"]))
            feat.append(np.asarray(f["feat"], np.float32))
        for c in ch:
            chip_crc.append(crc(c))
            if len(chips) < 10:
                chips.append(np.ascontiguousarray(c))
        state.append([R._no_face_streak, R._rot_cycle, R._prescan_rr, R._frame_idx, R._last_face_idx, n_blob])
        n_ex += 1
    out["ex_call_meta"] = np.array(call_meta, np.int64).reshape(-1, 6)
    out["ex_call_thresh"] = np.array(call_thresh, np.float64)
    out["ex_det"] = np.concatenate(dets) if dets else np.zeros((0, 5), np.float32)
    out["ex_kps"] = np.concatenate(kpss) if kpss else np.zeros((0, 5, 2), np.float32)
    out["ex_out_count"] = np.array(out_count, np.int32)
    out["ex_bbox"] = np.array(bbox, np.int32).reshape(-1, 4)
    out["ex_quality"] = np.array(quality, np.float64)
    out["ex_feat"] = np.array(feat, np.float32).reshape(-1, 512)
    out["ex_chip_crc"] = np.array(chip_crc, np.int64)
    out["ex_chips"] = np.stack(chips) if chips else np.zeros((0, 112, 112, 3), np.uint8)
    out["ex_state"] = np.array(state, np.int64)
    print(f"extract: {n_ex} calls of extract, {len(call_meta)} SCRFD passes, {len(bbox)} faces")


def make_processor(ga, cfg, fps, total_frames):
    P = object.__new__(ga.Processor)
    P.cfg = cfg
    P._cmd_q = queue.Queue()
    P._paused = False
    P._abort = False
    P._speed = 1.0
    P._prescan_cache_dirty = False
    P._hdr_preview_reader = None
    P._total_frames = int(total_frames)
    P._keyframes = []
    P._fps = float(fps)
    P.progress = RH._Inert()
    P.status = RH._Inert()
    P._status = lambda *a, **k: None
    P._emit_preview_bgr = lambda *a, **k: None
    P._hdr_preview_enabled = lambda *a, **k: False
    return P


def make_cfg(ga, overrides):
    cfg = ga.SessionConfig()
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def gen_bank(ga, out):
    P = make_processor(ga, make_cfg(ga, {}), 24, 100)
    for name, (cfg_over, feats, quals, seed_rows) in S.bank_cases().items():
        cfg = make_cfg(ga, cfg_over)
        bank_list = [r.copy() for r in seed_rows]
        bank = np.vstack(bank_list).astype(np.float32) if bank_list else None
        actions, idxs, fds = [], [], []
        for v, q in zip(feats, quals):
            fds.append(ga.Processor._fd_min(v, bank))
            bank, act, idx = P._stream_ref_bank_update(bank_list, bank, v, float(q), cfg)
            actions.append(S.ACTIONS.index(act))
            idxs.append(-1 if idx is None else int(idx))
        out[f"bank_{name}_actions"] = np.array(actions, np.int32)
        out[f"bank_{name}_idx"] = np.array(idxs, np.int32)
        out[f"bank_{name}_fd"] = np.array(fds, np.float64)
        out[f"bank_{name}_final"] = np.asarray(bank, np.float32)
    # _fd_min corner cases
    v = S.unit_vec(7)
    out["fdmin_corner"] = np.array([ga.Processor._fd_min(None, np.eye(4, 512, dtype=np.float32)), ga.Processor._fd_min(v, None),
                                    ga.Processor._fd_min(v, np.zeros((0, 512), np.float32)), ga.Processor._fd_min(v, v),
                                    ga.Processor._fd_min(3.0 * v, np.stack([S.unit_vec(8), v]))], np.float64)


def gen_prescan(ga, out):
    for name, case in S.prescan_cases().items():
        cfg = make_cfg(ga, case["cfg"])
        n, fps = case["n"], case["fps"]
        sc, ref = S.prescan_inputs(case)
        P = make_processor(ga, cfg, fps, n)
        face = S.RecordingFakeFace(sc)
        cap = S.PipeLikeCap(n, case.get("h", 2), case.get("w", 2))
        spans, bank = P._prescan(cap, fps, n, face, ref, cfg)
        out[f"ps_{name}_spans"] = np.asarray(spans, np.int64).reshape(-1, 2)
        out[f"ps_{name}_bank"] = np.zeros((0, 512), np.float32) if bank is None else np.asarray(bank, np.float32).reshape(-1, 512)
        out[f"ps_{name}_calls"] = np.array(face.calls, np.int64).reshape(-1, 4)      # (frame, rr_mode full?, escalate, frame width)
        out[f"ps_{name}_final_state"] = np.array([face.conf, float(face.rot_adaptive), float(face._prescan_escalate),
                                                  float(face._prescan_rr_mode == "full"), float(face._prescan_rr), float(face._frame_idx),
                                                  float(getattr(face, "_fast", -1))], np.float64)
        print(f"prescan[{name}]: {len(face.calls)} extract calls, spans {spans}, bank rows {0 if bank is None else len(bank)}")


def gen_prescan_full(fe, ga, out):
    """The reference's _prescan driving the reference's FaceEmbedder (both unmodified) over a rendered clip: the composition
    of the loop's knob changes with the embedder's rotation / flip state, frame downscale included."""
    rec = RecordingScrfd(SCRFDOracle(H.oracle_scrfd(S.EXTRACT_SCRFD)))
    R = RH.make_reference_embedder(fe, rec, H.proj_arcface, conf=0.5)
    frames, ref_img = S.full_clip_frames()
    cfg = make_cfg(ga, S.FULL_CFG)
    rfaces = R.extract(ref_img)
    ref = np.asarray(max(rfaces, key=lambda f: f["quality"])["feat"], np.float32)[None]
    P = make_processor(ga, cfg, S.FULL_FPS, S.FULL_N)
    spans, bank = P._prescan(S.FrameCap(frames), S.FULL_FPS, S.FULL_N, R, ref, cfg)
    out["pf_call_meta"] = np.array([[0, c["crc"], c["h"], c["w"], c["size"], len(c["det"])] for c in rec.calls], np.int64).reshape(-1, 6)
    out["pf_call_thresh"] = np.array([c["thresh"] for c in rec.calls], np.float64)
    out["pf_det"] = np.concatenate([c["det"] for c in rec.calls])
    out["pf_kps"] = np.concatenate([c["kps"] for c in rec.calls])
    out["pf_ref"] = ref
    out["pf_spans"] = np.asarray(spans, np.int64).reshape(-1, 2)
    out["pf_bank"] = np.asarray(bank, np.float32).reshape(-1, 512)
    out["pf_state"] = np.array([R._no_face_streak, R._rot_cycle, R._prescan_rr, R._frame_idx, R._last_face_idx, int(R._fast_prescan),
                                int(R._prescan_escalate), int(R.rot_adaptive)], np.int64)
    print(f"prescan_full: {len(rec.calls)} SCRFD passes, spans {spans}, bank rows {len(bank)}")


class DnnScrfdNet:
    """SCRFDOracle's `net.run(blob)` on cv2.dnn executing the ONNX export of the graph (one fixed-size file per S)."""

    def __init__(self, name, workdir):
        from person_capture_b200 import weights
        self.name, self.params, self.dir, self.nets = name, weights.load_params(name), workdir, {}

    def run(self, blob):
        from oracle import onnx_export as X
        S = int(blob.shape[2])
        if S not in self.nets:
            path = os.path.join(self.dir, f"{self.name}_{S}.onnx")
            names = X.export_scrfd(self.name, self.params, S, path, layout="insightface")
            self.nets[S] = (cv2.dnn.readNetFromONNX(path), names)
        net, names = self.nets[S]
        net.setInput(np.ascontiguousarray(blob, np.float32))
        return [np.asarray(o).copy() for o in net.forward(names)]


def gen_reference_onnx(fe, out):
    """The unmodified reference FaceEmbedder whose two sessions execute the EXPORTED ONNX GRAPHS through cv2.dnn (an engine
    independent of torch and of our oracle's executors): the offline stand-in for "the reference's ONNX Runtime CPU run of the
    same graphs".  The GPU tests hold the CUDA path to these vectors directly -- no oracle in between."""
    from oracle import onnx_export as X
    from person_capture_b200 import weights
    with tempfile.TemporaryDirectory() as td:
        arc_path = os.path.join(td, "arc.onnx")
        arc_out = X.export_iresnet(S.RO_ARC, weights.load_params(S.RO_ARC), arc_path)
        arc_net = cv2.dnn.readNetFromONNX(arc_path)

        def arc_fn(x):
            rows = []
            for k in range(x.shape[0]):
                arc_net.setInput(np.ascontiguousarray(x[k:k + 1], np.float32))
                rows.append(np.asarray(arc_net.forward(arc_out)).reshape(512).copy())
            return np.stack(rows).astype(np.float32)

        R = RH.make_reference_embedder(fe, SCRFDOracle(DnnScrfdNet(S.RO_SCRFD, td)), arc_fn, conf=0.5)
        R.configure_rotation_strategy(adaptive=False)
        R.set_prescan_fast(True, mode="rr")
        R._prescan_probe_imgsz = 512
        counts, bbox, quality, feat = [], [], [], []
        for k, key in enumerate(S.RO_FRAMES):
            R.set_prescan_hint(escalate=bool(k % 2))
            faces = R.extract(S.ro_frame(key))
            counts.append(len(faces))
            for f in faces:
                bbox.append(np.asarray(f["bbox"], np.int32))
                quality.append(float(f["quality"]))
                feat.append(np.asarray(f["feat"], np.float32))
        out["ro_counts"] = np.array(counts, np.int32)
        out["ro_bbox"] = np.array(bbox, np.int32).reshape(-1, 4)
        out["ro_quality"] = np.array(quality, np.float64)
        out["ro_feat"] = np.array(feat, np.float32).reshape(-1, 512)
        out["ro_state"] = np.array([R._prescan_rr, R._no_face_streak, R._frame_idx], np.int64)
        print(f"reference x onnx: {counts} faces per frame")

        # the bench's models on the bench's frame shape: SCRFD-10G + iResNet-100, 960x540
        arc100_path = os.path.join(td, "arc100.onnx")
        arc100_out = X.export_iresnet(S.RH_ARC, weights.load_params(S.RH_ARC), arc100_path)
        arc100 = cv2.dnn.readNetFromONNX(arc100_path)

        def arc100_fn(x):
            rows = []
            for k in range(x.shape[0]):
                arc100.setInput(np.ascontiguousarray(x[k:k + 1], np.float32))
                rows.append(np.asarray(arc100.forward(arc100_out)).reshape(512).copy())
            return np.stack(rows).astype(np.float32)

        RH_ = RH.make_reference_embedder(fe, SCRFDOracle(DnnScrfdNet(S.RH_SCRFD, td)), arc100_fn, conf=0.5)
        RH_.configure_rotation_strategy(adaptive=False)
        RH_.set_prescan_fast(True, mode="rr")
        RH_._prescan_probe_imgsz = 512
        counts, bbox, quality, feat = [], [], [], []
        for k, i in enumerate(S.RH_FRAME_IDS):
            RH_.set_prescan_hint(escalate=bool(k % 2))
            faces = RH_.extract(S.rh_frame(i))
            counts.append(len(faces))
            for f in faces:
                bbox.append(np.asarray(f["bbox"], np.int32))
                quality.append(float(f["quality"]))
                feat.append(np.asarray(f["feat"], np.float32))
        out["rh_counts"] = np.array(counts, np.int32)
        out["rh_bbox"] = np.array(bbox, np.int32).reshape(-1, 4)
        out["rh_quality"] = np.array(quality, np.float64)
        out["rh_feat"] = np.array(feat, np.float32).reshape(-1, 512)
        out["rh_state"] = np.array([RH_._prescan_rr, RH_._no_face_streak, RH_._frame_idx], np.int64)
        print(f"reference x onnx (10G + R100, 960x540): {counts} faces per frame")

        # and the whole pre-scan: Processor._prescan x FaceEmbedder x ONNX graphs, reference bank built the reference's way
        # (gui_app.py:4517-4556: each reference image and its mirror, best face, streaming update)
        import importlib
        ga = importlib.import_module("person_capture.gui_app")
        R2 = RH.make_reference_embedder(fe, SCRFDOracle(DnnScrfdNet(S.RO_SCRFD, td)), arc_fn, conf=0.5)
        frames, ref_img = S.rp_clip_frames()
        cfg = make_cfg(ga, S.RP_CFG)
        R2.conf = float(cfg.face_det_conf)
        P = make_processor(ga, cfg, S.RP_FPS, S.RP_N)
        bank_list, ref = [], None
        for aug in (ref_img, cv2.flip(ref_img, 1)):
            bf = fe.FaceEmbedder.best_face(R2.extract(aug))
            if bf and bf.get("feat") is not None:
                ref, _, _ = P._stream_ref_bank_update(bank_list, ref, bf["feat"], float(bf.get("quality", 0.0)), cfg)
        n_ext = [0]
        inner = R2.extract

        def counted(img, **kw):
            n_ext[0] += 1
            return inner(img, **kw)

        R2.extract = counted
        spans, bank = P._prescan(S.FrameCap(frames), S.RP_FPS, S.RP_N, R2, ref, cfg)
        out["rp_ref"] = np.asarray(ref, np.float32).reshape(-1, 512)
        out["rp_spans"] = np.asarray(spans, np.int64).reshape(-1, 2)
        out["rp_bank"] = np.asarray(bank, np.float32).reshape(-1, 512)
        out["rp_extracts"] = np.array([n_ext[0]], np.int64)
        print(f"reference x onnx pre-scan: {n_ext[0]} extract calls, spans {spans}, bank {len(ref)} -> {len(bank)} rows")

        # the same with the bench's models on a 1080p clip (pre-scan downscale to 960 wide inside the reference's loop)
        R3 = RH.make_reference_embedder(fe, SCRFDOracle(DnnScrfdNet(S.RH_SCRFD, td)), arc100_fn, conf=0.5)
        frames, ref_img = S.rq_clip_frames()
        cfg = make_cfg(ga, S.RQ_CFG)
        R3.conf = float(cfg.face_det_conf)
        P = make_processor(ga, cfg, S.RQ_FPS, S.RQ_N)
        bank_list, ref = [], None
        for aug in (ref_img, cv2.flip(ref_img, 1)):
            bf = fe.FaceEmbedder.best_face(R3.extract(aug))
            if bf and bf.get("feat") is not None:
                ref, _, _ = P._stream_ref_bank_update(bank_list, ref, bf["feat"], float(bf.get("quality", 0.0)), cfg)
        spans, bank = P._prescan(S.FrameCap(frames), S.RQ_FPS, S.RQ_N, R3, ref, cfg)
        out["rq_ref"] = np.asarray(ref, np.float32).reshape(-1, 512)
        out["rq_spans"] = np.asarray(spans, np.int64).reshape(-1, 2)
        out["rq_bank"] = np.asarray(bank, np.float32).reshape(-1, 512)
        print(f"reference x onnx pre-scan (10G + R100, 1080p): spans {spans}, bank {len(ref)} -> {len(bank)} rows")


def gen_cache(ga, out):
    import shutil
    video, ref = S.cache_files()
    cfg = make_cfg(ga, dict(S.CACHE_CFG, video=video, ref=ref, prescan_cache_dir=os.path.join(S.CACHE_DIR, "cache")))
    P = make_processor(ga, cfg, 23.976, 4321)
    meta = P._prescan_cache_meta(cfg, 23.976, 4321)
    path = P._prescan_cache_path(cfg, meta)
    spans = [(10, 200), (400, 4320)]
    bank = np.stack([S.unit_vec(i) for i in range(3)])
    P._save_prescan_cache(cfg, 23.976, 4321, spans, bank)
    with np.load(str(path), allow_pickle=False) as z:
        arrays = {k: z[k] for k in z.files}
    out["cache_key_json"] = np.array(json.dumps({k: v for k, v in meta.items() if k != "key"}, sort_keys=True, separators=(",", ":")))
    out["cache_key"] = np.array(str(meta["key"]))
    out["cache_path_name"] = np.array(os.path.basename(str(path)))
    out["cache_file_keys"] = np.array(sorted(arrays.keys()))
    out["cache_file_meta"] = np.array(str(arrays["meta"].item()))
    out["cache_file_spans"] = arrays["spans"]
    out["cache_file_ref"] = arrays["ref_face_feat"]
    out["cache_file_has_ref"] = arrays["has_ref"]
    out["cache_file_dtypes"] = np.array([str(arrays[k].dtype) for k in sorted(arrays.keys())])
    hit, lspans, lref, _ = P._load_prescan_cache(cfg, 23.976, 4321)
    assert hit and [tuple(x) for x in lspans] == spans and np.array_equal(lref, bank)
    # and without a reference bank
    cfg2 = make_cfg(ga, dict(S.CACHE_CFG, video=video, ref="", prescan_cache_dir=os.path.join(S.CACHE_DIR, "cache"), prescan_stride=7))
    meta2 = P._prescan_cache_meta(cfg2, 30.0, 100)
    P._save_prescan_cache(cfg2, 30.0, 100, [], None)
    with np.load(str(P._prescan_cache_path(cfg2, meta2)), allow_pickle=False) as z:
        out["cache2_ref_shape"] = np.array(z["ref_face_feat"].shape, np.int64)
        out["cache2_has_ref"] = z["has_ref"]
        out["cache2_spans_shape"] = np.array(z["spans"].shape, np.int64)
    out["cache2_key"] = np.array(str(meta2["key"]))
    # raw bytes of the first file, so that a test can feed the reference's own file to the product's loader
    out["cache_file_bytes"] = np.frombuffer(open(str(path), "rb").read(), np.uint8)
    shutil.rmtree(S.CACHE_DIR, ignore_errors=True)


def gen_geometry(ga, out):
    cases, pairs = S.geometry_cases()
    P = make_processor(ga, make_cfg(ga, {}), 24, 10)
    out["geo_expand"] = np.array([ga.Processor._expand_xyxy(b, px, py, W, Hh) for b, px, py, W, Hh in cases], np.int64)
    out["geo_clip"] = np.array([ga.Processor._clip_to_frame(b[0], b[1], b[2], b[3], W, Hh) for b, px, py, W, Hh in cases], np.int64)
    out["geo_iou"] = np.array([P._iou(a, b) for a, b in pairs], np.float64)


def gen_curator(out):
    import importlib
    dc = importlib.import_module("person_capture.dataset_curator")
    rows = []
    canv_crc = []
    for square in (True, False):
        for passed in (False, True):
            C = object.__new__(dc.Curator)
            C.face = S.CannedFace()
            C._det_square = square
            C._kps_enabled = False
            C._progress = None
            C.id_already_passed = passed
            C.ref_feat = S.unit_vec(3) * np.float32(1.3)
            for img in S.curator_images():
                best = C._detect_best_face(img)
                if best is None:
                    rows.append([0, 0, 0, 0, 0, 0.0, C._fd_min(None)])
                else:
                    rows.append([1, *[int(v) for v in best["bbox"]], float(best["quality"]), C._fd_min(best["feat"])])
            canv_crc.extend([c for (_, _, c) in C.face.seen])
    out["cur_rows"] = np.array(rows, np.float64)
    out["cur_canvas_crc"] = np.array(canv_crc, np.int64)
    lb = [dc.Curator._letterbox_square(img, 640) for img in S.curator_images()]
    out["cur_lb_meta"] = np.array([[sc, dx, dy] for (_, sc, dx, dy) in lb], np.float64)
    out["cur_lb_crc"] = np.array([crc(cv) for (cv, _, _, _) in lb], np.int64)
    C.ref_feat = None
    out["cur_fd_noref"] = np.array([C._fd_min(S.unit_vec(1))], np.float64)


def main():
    fe, ga = RH.import_reference()
    out = {}
    gen_curator(out)
    gen_geometry(ga, out)
    gen_units(fe, out)
    gen_bank(ga, out)
    gen_prescan(ga, out)
    gen_cache(ga, out)
    gen_prescan_full(fe, ga, out)
    gen_reference_onnx(fe, out)
    gen_extract(fe, out)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
