"""Generates tests/golden/oracle_golden.npz: outputs of the CPU oracle (torch-CPU fp32 + real cv2 4.13) on
fixed seeded inputs.  These pin the *oracle including its networks* across machines and over time (the reference's own
non-network code is pinned by make_reference_golden.py / reference_golden.npz); the cv2-level
known answers (LMedS subset sequence, fixed-point resize/warp) are pinned against cv2 itself in
tests/test_cpu_arith.py.   usage: python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pcb_test_helpers as H  # noqa: E402
from oracle import cv_emul, prescan as OP  # noqa: E402
from oracle.face_embedder import arcface_preprocess  # noqa: E402
from person_capture_b200 import synth  # noqa: E402
from person_capture_b200.params import PrescanParams  # noqa: E402


def build():
    out = {}
    out["lmeds_pairs_5"] = np.array(cv_emul.lmeds_pairs(5), np.int32)
    out["lmeds_pairs_3"] = np.array(cv_emul.lmeds_pairs(3), np.int32)
    clip = synth.ClipSpec(416, 234, 60, seed=1001, target_segments=[(0, 59)])
    frame = clip.frame(11)
    out["frame_crc"] = np.array([int(frame.astype(np.int64).sum()), int((frame.astype(np.int64) ** 2).sum())], np.int64)
    face = H.oracle_embedder("scrfd_2.5g_bnkps", "arcface_r50", conf=0.5)
    face.scrfd.det_thresh = 0.5
    det, kps = face.scrfd.detect(frame, (416, 416))
    out["det"], out["kps"] = det.astype(np.float32), kps.astype(np.float32)
    faces = face.extract(frame)
    out["ex_bbox"] = np.stack([f["bbox"] for f in faces]).astype(np.int32)
    out["ex_quality"] = np.array([f["quality"] for f in faces], np.float64)
    out["ex_feat"] = np.stack([f["feat"] for f in faces]).astype(np.float32)
    chip = face.last_chips[0]
    out["chip"] = chip
    out["emb_raw16"] = H.oracle_arcface("arcface_r50").run(arcface_preprocess(chip)[None])[0, :16].astype(np.float32)
    # a small pre-scan (config-1 shape: 640x360 -> 416x234, SCRFD-2.5G + R50)
    cfg = PrescanParams(face_model="scrfd_2.5g_bnkps", prescan_stride=6, prescan_max_width=416, prescan_fd_enter=0.62,
                        prescan_fd_exit=0.72, prescan_fd_add=0.50, face_quality_min=40.0, prescan_min_segment_sec=0.5,
                        prescan_pad_sec=0.25, prescan_bridge_gap_sec=0.25, prescan_exit_cooldown_sec=0.25,
                        prescan_boundary_refine_sec=0.5)
    c2 = synth.ClipSpec(640, 360, 144, seed=1001)
    frames = [c2.frame(i) for i in range(144)]
    ora = H.oracle_embedder("scrfd_2.5g_bnkps", "arcface_r50", conf=cfg.face_det_conf)
    bank = OP.build_reference_bank(ora, [synth.reference_image(1, 512, seed=1001)], cfg)
    log = []
    spans, bank2 = OP.prescan(lambda i: frames[i] if i < 144 else None, 24, 144, ora, bank, cfg, log=log)
    out["ps_spans"] = np.asarray(spans, np.int64).reshape(-1, 2)
    out["ps_best"] = np.array([r["best"] for r in log], np.float64)
    out["ps_skip"] = np.array([r["skip"] for r in log], np.uint8)
    out["ps_bank_rows"] = np.array([bank.shape[0], np.asarray(bank2).shape[0]], np.int32)
    return out


if __name__ == "__main__":
    g = build()
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **g)
    print({k: v.shape for k, v in g.items()})
    print("spans", g["ps_spans"].tolist(), "bank", g["ps_bank_rows"].tolist())
