"""GPU timeline of the host-resident pre-scan step (bench.py's e2e leg): PCB_TIMELINE events of the copy stream, the main
stream and the flip context.  usage: PCB_TIMELINE=gpurun_out/tl.jsonl python tools/e2e_timeline.py ; prints the last step."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = os.environ.setdefault("PCB_TIMELINE", "gpurun_out/timeline.jsonl")
if os.path.exists(path):
    os.remove(path)
import torch
import bench
from person_capture_b200 import prescan as PS
from person_capture_b200.face_embedder import FaceEmbedder
cfg = bench.make_cfg()
face = FaceEmbedder("cuda:0", "scrfd_10g_bnkps", conf=cfg.face_det_conf, arcface_model="arcface_r100")
frames, ref = bench.make_pool(64)
bank = PS.build_reference_bank(face, [ref], cfg)
clip = bench.PooledHostClip(torch.from_numpy(frames).pin_memory(), 512)
for i in range(3):
    stats = {}
    PS.prescan_batched(clip, 24, face, bank, cfg, batch=64, stats=stats)
    print("step", i, stats["phase_ms"], "early rows", stats.get("early_flip_rows"))
last = json.loads(open(path).read().strip().splitlines()[-1])
for name, t in last:
    print(f"{t:9.3f}  {name}")
