"""Config 3 at N GPUs: SCRFD-10G + ArcFace R100 main pass over kept spans of a 4K clip (face_fullframe_imgsz=1280), spans
sharded across the ranks (mainpass.main_pass_sharded), every rank ending with ALL hits.  One JSON line from rank 0:

  python tools/bench_mainpass.py                                            # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 tools/bench_mainpass.py

Weak scaling: every rank brings `--spans-per-rank` spans of `--span-len` frames (frame_stride 2), so the clip grows with N.
Frames are synthetic 4K frames resident in HBM (a pool cycled over the clip); the sharded result is checked against the
sequential main_pass on rank 0 (outside the timed region).  Timing: CUDA events on the engine's stream, max over ranks."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--spans-per-rank", type=int, default=4)
ap.add_argument("--span-len", type=int, default=96)
ap.add_argument("--gap", type=int, default=24)
ap.add_argument("--pool", type=int, default=12)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--no-check", action="store_true")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if world > 1 else 0

from person_capture_b200 import mainpass as MP, prescan as PS, synth  # noqa: E402
from person_capture_b200.face_embedder import FaceEmbedder  # noqa: E402
from person_capture_b200.params import PrescanParams  # noqa: E402

cfg = PrescanParams(face_model="scrfd_10g_bnkps", face_thresh=0.62, face_quality_min=40.0, face_fullframe_imgsz=1280,
                    frame_stride=2, face_fullframe_cadence=6, lock_face_roi_max_misses=3)
face = FaceEmbedder(f"cuda:{local}", "scrfd_10g_bnkps", conf=cfg.face_det_conf, arcface_model="arcface_r100")
eng = face.engine
bank = PS.build_reference_bank(face, [synth.reference_image(1, 512, seed=2101)], cfg)

n_spans = args.spans_per_rank * world
total = n_spans * (args.span_len + args.gap)
spans = [(k * (args.span_len + args.gap) + args.gap // 2, k * (args.span_len + args.gap) + args.gap // 2 + args.span_len - 1) for k in range(n_spans)]
spec = synth.ClipSpec(3840, 2160, max(args.pool, 2), seed=2101, target_segments=[(0, args.pool)], face_px=(60, 110))
pool = eng.to_device(np.stack([spec.frame(i) for i in range(args.pool)]))


class PooledClip:
    """`total` frames backed by a pool of distinct frames in HBM (frame i = pool[i % pool])."""

    def __init__(self, frames, n):
        self.pool = frames
        self.total_frames = int(n)

    def host(self, i):
        return self.pool[i % self.pool.shape[0]].cpu().numpy()

    def device_batch(self, eng_, idxs, stream=None):
        P = self.pool.shape[0]
        with torch.cuda.stream(eng_.stream):
            sel = torch.as_tensor([i % P for i in idxs], device=self.pool.device)
            return self.pool.index_select(0, sel)


clip = PooledClip(pool, total)


def run():
    st = {}
    hits = MP.main_pass_sharded(clip, 24.0, spans, face, bank, cfg, stats=st)
    return hits, st


for _ in range(args.warmup):
    run()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(eng.stream)
t0 = time.perf_counter()
for _ in range(args.steps):
    hits, st = run()
e1.record(eng.stream)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
ms = e0.elapsed_time(e1)
if world > 1:
    t = torch.tensor([ms], device=eng.tdev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
frames_eval = sum(len(range(s, e + 1, cfg.frame_stride)) for s, e in spans)
equal = None
if not args.no_check and rank == 0:
    face2 = FaceEmbedder(f"cuda:{local}", "scrfd_10g_bnkps", conf=cfg.face_det_conf, engine=eng, arcface_model="arcface_r100")
    ref = MP.main_pass(clip, 24.0, spans, face2, bank, cfg)
    key = lambda h: (h["idx"], h["site"], tuple(h["face_box"]), round(h["fd"], 6))
    equal = [key(h) for h in hits] == [key(h) for h in ref]
    if not equal:
        print(f"sharded {len(hits)} hits vs sequential {len(ref)}", file=sys.stderr)
        for a, b in zip(hits, ref):
            if key(a) != key(b):
                print("first difference:", key(a), "vs", key(b), file=sys.stderr)
                break
        # a second sequential run on yet another fresh embedder: is the sequential pass itself repeatable?
        face3 = FaceEmbedder(f"cuda:{local}", "scrfd_10g_bnkps", conf=cfg.face_det_conf, engine=eng, arcface_model="arcface_r100")
        ref2 = MP.main_pass(clip, 24.0, spans, face3, bank, cfg)
        print("sequential repeatable:", [key(h) for h in ref2] == [key(h) for h in ref], file=sys.stderr)
if world > 1:
    dist.barrier()
if rank == 0:
    print(json.dumps({"metric": "main_pass_frames_per_sec", "value": frames_eval * args.steps / (ms / 1000.0), "unit": "frames/s", "n_gpus": world,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                      "dtype": "f16", "data": "synthetic",
                      "config": {"workload": "SCRFD-10G + ArcFace R100 main pass, 4K frames, face_fullframe_imgsz=1280, frame_stride 2, "
                                             "kept spans sharded in contiguous blocks", "spans": n_spans, "span_len": args.span_len,
                                 "frames_evaluated_per_step": frames_eval, "pool_frames": args.pool},
                      "hits": len(hits), "sites": sorted({h["site"] for h in hits}), "fixup_rounds": st.get("rounds"),
                      "equals_sequential_main_pass": equal, "wall_ms_per_step": 1000.0 * wall / args.steps}))
if world > 1:
    dist.destroy_process_group()
