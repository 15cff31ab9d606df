"""Distribution of GPU-vs-oracle differences on the small config (2.5G + R50): landmarks, quality, fd."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcb_test_helpers as H
from oracle import prescan as OP
from person_capture_b200 import synth, prescan as PS
from person_capture_b200.params import PrescanParams
from person_capture_b200.face_embedder import FaceEmbedder
cfg = PrescanParams(face_model="scrfd_2.5g_bnkps")
face = FaceEmbedder("cuda:0", "scrfd_2.5g_bnkps", conf=0.5, arcface_model="arcface_r50")
ora = H.oracle_embedder("scrfd_2.5g_bnkps", "arcface_r50", conf=0.5)
ref = synth.reference_image(1, 512, seed=1001)
gb, ob = PS.build_reference_bank(face, [ref], cfg), OP.build_reference_bank(ora, [ref], cfg)
print("bank cos", [round(H.cos(a, b), 6) for a, b in zip(gb, ob)])
for f in (face, ora):
    f.configure_rotation_strategy(adaptive=False); f.set_prescan_fast(True); f._prescan_probe_imgsz = 512
clip = synth.ClipSpec(416, 234, 300, seed=1001)
dfd, dq, dcos, dfd_samechip = [], [], [], []
for i in range(0, 300, 5):
    fr = clip.frame(i)
    g, o = face.extract(fr), ora.extract(fr)
    if len(g) != len(o) or not g:
        continue
    # embedding of the ORACLE's chips on the GPU isolates ArcFace numerics from chip differences
    chips = np.stack(ora.last_chips)
    import torch
    e, _ = face.engine.embed(face.engine.to_device(chips), len(chips), False)
    face.engine.sync()
    e = e.cpu().numpy(); e /= np.linalg.norm(e, axis=1, keepdims=True)
    # oracle order: last_chips is in kept order; recompute oracle feats in that order
    of = ora.arcface_encode(ora.last_chips)
    for a, b in zip(e, of):
        dfd_samechip.append(abs(OP.fd_min(a, ob) - OP.fd_min(b, ob)))
    for a, b in zip(g, o):
        dfd.append(abs(OP.fd_min(a["feat"], gb) - OP.fd_min(b["feat"], ob)))
        dq.append(abs(a["quality"] - b["quality"]) / max(1.0, b["quality"]))
        dcos.append(1 - H.cos(a["feat"], b["feat"]))
for name, v in (("|dfd| e2e", dfd), ("|dfd| same chips (ArcFace numerics only)", dfd_samechip), ("rel dquality", dq), ("1-cos", dcos)):
    v = np.array(v)
    print(f"{name:45s} n={len(v)} mean={v.mean():.2e} p50={np.median(v):.2e} p95={np.percentile(v,95):.2e} max={v.max():.2e}")
