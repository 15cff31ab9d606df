"""Diagnose e2e outliers: frames where GPU and oracle boxes differ; compares chips and alignment kinds."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcb_test_helpers as H
from person_capture_b200 import synth
from person_capture_b200.face_embedder import FaceEmbedder
face = FaceEmbedder("cuda:0", "scrfd_10g_bnkps", conf=0.5, arcface_model="arcface_r50")
face.keep_debug = True
ora = H.oracle_embedder("scrfd_10g_bnkps", "arcface_r50", conf=0.5)
for f in (face, ora):
    f.configure_rotation_strategy(adaptive=False); f.set_prescan_fast(True, mode="rr"); f._prescan_probe_imgsz = 512
clip = synth.ClipSpec(960, 540, 120, seed=77)
for i in range(0, 120, 9):
    frame = clip.frame(i)
    if i % 2:
        for f in (face, ora): f.set_prescan_hint(escalate=True)
    got, ref = face.extract(frame), ora.extract(frame)
    for f in (face, ora): f.set_prescan_hint(escalate=False)
    gch = face.last_chips[face.last_order] if got else []
    och = ora.last_chips if ref else []
    for k, (g, r) in enumerate(zip(got, ref)):
        c = H.cos(g["feat"], r["feat"])
        d = np.abs(g["bbox"].astype(int) - r["bbox"].astype(int)).max()
        if d or c < 0.999:
            extra = ""
            try:
                # oracle chips are in kept order; returned faces are sorted by (quality, area)
                extra = f" q_gpu={g['quality']:.1f} q_ora={r['quality']:.1f}"
            except Exception:
                pass
            print(f"frame {i} face {k}: dbox={d} cos={c:.4f} gbox={g['bbox'].tolist()} rbox={r['bbox'].tolist()}{extra}")
print("kinds(last)", getattr(face, "last_kinds", None))
