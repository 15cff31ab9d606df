"""Every kernel of libpcb200 once, at reduced sizes, for compute-sanitizer (tools/sanitize.sh): K0 resize (three modes), K1
letterbox (rotation + pad), the SCRFD-2.5G graph at S=320 (halo / strided / resident / streamed conv variants, pooling,
upsample-add), K3 decode + NMS, K4 align (all kinds), the iResNet-50 graph on 6 chips + flips (pair and two-issuer variants,
fp32 FC), K5 match, the live distance table (full + incremental refresh) through the native replay."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from person_capture_b200 import synth, prescan as PS
from person_capture_b200.engine import Engine
from person_capture_b200.face_embedder import FaceEmbedder
from person_capture_b200.params import PrescanParams

eng = Engine(0, scrfd="scrfd_2.5g_bnkps", arcface="arcface_r50")
rng = np.random.default_rng(0)
clip = synth.ClipSpec(416, 234, 24, seed=1001, target_segments=[(4, 20)])
frames = np.stack([clip.frame(i) for i in range(0, 24, 2)])
dev = eng.to_device(frames)
big = eng.to_device(rng.integers(0, 256, (2, 468, 832, 3), dtype=np.uint8))
eng.resize(big, 234, 416, area=True)
eng.resize(big, 200, 300, area=True)
eng.resize(big, 300, 500, area=False)
for rot, pad in ((0, 0), (90, 0), (270, 24), (180, 24)):
    eng.letterbox(dev[:2], 320, rot=rot, pad=pad, want_det_img=True)
det = eng.detect(dev, 320, 0.5)
al = eng.align(dev, det, max_faces=64)
eng.sync()
n = int(al.face_total.cpu()[0])
print("faces", n)
emb, embf = eng.embed(al.chips, max(n, 1), True)
eng.set_bank(rng.normal(size=(5, 512)).astype(np.float32))
eng.match(emb, embf, None, max(n, 1))
eng.sync()
cfg = PrescanParams(face_model="scrfd_2.5g_bnkps", prescan_stride=1, prescan_max_width=416, prescan_fd_enter=0.62, prescan_fd_exit=0.72,
                    prescan_fd_add=0.50, face_quality_min=40.0, prescan_min_segment_sec=0.25, prescan_pad_sec=0.1,
                    prescan_add_cooldown_samples=1, prescan_boundary_refine_sec=0.1)
face = FaceEmbedder("cuda:0", "scrfd_2.5g_bnkps", conf=0.5, engine=eng)
bank = PS.build_reference_bank(face, [synth.reference_image(1, 256, seed=1001)], cfg)
spans, bank2 = PS.prescan_batched(PS.DeviceClip(dev), 24, face, bank, cfg, batch=8)
eng.sync()
print("spans", spans, "bank rows", None if bank2 is None else np.asarray(bank2).shape[0], "launches", eng.launch_count())
eng.close()
print("SANITIZE_RUN_OK")
