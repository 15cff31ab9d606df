#!/usr/bin/env python
"""Benchmark of the identity hot path (BASELINE.json metric: pre-scan frames/sec, faces embedded/sec).

Workload (N=1): BASELINE.json configs[1] -- SCRFD-10G + ArcFace R100 pre-scan over 1080p frames with
prescan_max_width=960, stride 1.  One *step* = one complete pre-scan (K0 downscale -> K1 letterbox ->
SCRFD -> K3 decode/NMS -> K4 align -> ArcFace (+flip superset) -> K5 match -> span state machine ->
edge refinement) of a clip of --frames-per-step frames per GPU.  The clip's content cycles through a
pool of distinct synthetic frames (398 MB for 64 x 1080p, larger than the 126 MB L2).

  value      frames/s, whole job, frames resident in HBM when the timed region starts
  e2e        same metric through the public API with frames in pinned HOST memory (H2D inside)
  roofline   conv_tc2_kernel (tcgen05 implicit GEMM): algorithmic FLOPs / CUDA-event time of its launches
  cpu_baseline  the CPU oracle (torch-CPU fp32 + cv2, restated reference) on a bounded sample

`--impl reference` times the CPU oracle alone (the reference's own implementation cannot be installed
offline: ONNX Runtime / insightface / TensorRT are absent).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "SCRFD-10G + ArcFace R100 pre-scan, 1080p frames at prescan_max_width=960, stride 1"
METRIC = "prescan_frames_per_sec"


def make_cfg():
    from person_capture_b200.params import PrescanParams
    # thresholds calibrated on the synthetic identities (seeded-random ArcFace weights put same-identity distances
    # around 0.3-0.6, tests/test_gpu_e2e.py uses the same values): the clip then has active stretches, bank growth,
    # flip-TTA while active and edge refinement, as a real pre-scan does
    return PrescanParams(face_model="scrfd_10g_bnkps", prescan_stride=1, prescan_max_width=960, prescan_decode_max_w=0,
                         prescan_cache_mode="off", prescan_fd_enter=0.62, prescan_fd_exit=0.72, prescan_fd_add=0.50,
                         face_quality_min=40.0)


def make_pool(n: int, seed: int = 1002, distractor_prob: float = 1.0):
    """`n` distinct synthetic 1080p frames + reference image.  SURVEY.md 8(d) C2: about one face per frame -- a distractor
    identity in every frame, the target identity on top of it in two stretches (48 % of the frames).  distractor_prob < 1
    (secondary measurement) leaves frames without any face: rotated probes, heavy passes and the fd9 gate then carry load."""
    from person_capture_b200 import synth
    clip = synth.ClipSpec(1920, 1080, n, seed=seed, target=1, others=(2, 3, 4), distractor_prob=float(distractor_prob))
    frames = np.stack([clip.frame(i) for i in range(n)])
    return frames, synth.reference_image(1, 512, seed=seed)


class PooledDeviceClip:
    """Clip whose frame i is pool[i % P]; the pool is resident in HBM."""

    def __init__(self, pool, total_frames: int):
        self.pool = pool
        self.total_frames = int(total_frames)

    def host(self, i):
        return self.pool[i % self.pool.shape[0]].cpu().numpy()

    def device_batch(self, eng, idxs, stream=None):
        import torch
        P = self.pool.shape[0]
        lo = idxs[0] % P
        if lo + len(idxs) <= P and list(idxs) == list(range(idxs[0], idxs[0] + len(idxs))):
            return self.pool[lo:lo + len(idxs)]
        with torch.cuda.stream(eng.stream):
            sel = torch.as_tensor([i % P for i in idxs], device=self.pool.device)
            return self.pool.index_select(0, sel)


class PooledHostClip:
    """Same clip with the pool in pinned host memory: every batch is copied H2D inside the step."""

    def __init__(self, pool_pinned, total_frames: int):
        self.pool = pool_pinned
        self.total_frames = int(total_frames)
        self.h2d_bytes = 0

    def host(self, i):
        return self.pool[i % self.pool.shape[0]].numpy()

    host_resident = True

    def device_batch(self, eng, idxs, stream=None):
        import torch
        P = self.pool.shape[0]
        lo = idxs[0] % P
        if lo + len(idxs) <= P and list(idxs) == list(range(idxs[0], idxs[0] + len(idxs))):
            src = self.pool[lo:lo + len(idxs)]
        else:
            src = self.pool[torch.as_tensor([i % P for i in idxs])].pin_memory()
        self.h2d_bytes += src.numel()
        with torch.cuda.stream(stream if stream is not None else eng.stream):
            return src.to(eng.tdev, non_blocking=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def spans_digest(spans, bank) -> str:
    """sha1 over the kept spans and the shape of the final bank: what must be identical on every rank and at every N."""
    import hashlib
    rows = 0 if bank is None else int(np.asarray(bank).reshape(-1, 512).shape[0])
    return hashlib.sha1(json.dumps([[int(a), int(b)] for a, b in spans] + [rows]).encode()).hexdigest()[:16]


def cpu_oracle_rate(frames_u8, ref_img, cfg, sample: int, threads: int, out: dict | None = None):
    """Frames/s of the CPU oracle pre-scan loop over the first `sample` frames (bounded).  `out` receives the oracle's
    per-sample log, spans and bank (the parity block of the bench line compares the GPU path with them)."""
    import torch
    torch.set_num_threads(threads)
    import cv2
    cv2.setNumThreads(threads)
    from oracle import prescan as OP
    from oracle.face_embedder import FaceEmbedderOracle
    from oracle.models import FoldedIResNet, FoldedSCRFD
    from oracle.scrfd_detect import SCRFDOracle
    from person_capture_b200 import weights
    face = FaceEmbedderOracle(SCRFDOracle(FoldedSCRFD("scrfd_10g_bnkps", weights.load_params("scrfd_10g_bnkps"))),
                              FoldedIResNet("arcface_r100", weights.load_params("arcface_r100")), conf=cfg.face_det_conf)
    bank = OP.build_reference_bank(face, [ref_img], cfg)
    P = len(frames_u8)
    OP.prescan(lambda i: frames_u8[i % P] if i < 4 else None, 24, 4, face, bank, cfg)     # untimed: thread pools, first-touch pages
    log = []
    t0 = time.perf_counter()
    spans, bank2 = OP.prescan(lambda i: frames_u8[i % P] if i < sample else None, 24, sample, face, bank, cfg, log=log)
    dt = time.perf_counter() - t0
    if out is not None:
        out.update(log=log, spans=spans, bank=bank2, bank0=bank)
    return sample / dt, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = make_cfg()
    threads = os.cpu_count() or 1
    sample = args.ref_sample
    frames, ref_img = make_pool(min(sample, args.pool), distractor_prob=args.distractor_prob)
    rates = []
    for s in range(args.warmup + args.steps):
        r, _ = cpu_oracle_rate(frames, ref_img, cfg, sample, threads)
        if s >= args.warmup:
            rates.append(r)
    value = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * sample / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": sample},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} frames of the same clip per step through the CPU oracle "
                                       "(torch-CPU fp32 + cv2; ONNX Runtime is not installable offline)"},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=512, help="clip length per GPU per step")
    ap.add_argument("--pool", type=int, default=64, help="distinct 1080p frames resident in HBM")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--cpu-sample", type=int, default=96, help="frames of the clip the CPU oracle is timed on (cpu_baseline + parity block)")
    ap.add_argument("--ref-sample", type=int, default=48, help="frames per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--distractor-prob", type=float, default=1.0,
                    help="share of 24-frame blocks with a non-target face (default 1.0 = the headline workload; lower: frames without faces)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from person_capture_b200.face_embedder import FaceEmbedder
    from person_capture_b200 import prescan as PS, graphs, _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = make_cfg()
    face = FaceEmbedder(f"cuda:{local}", "scrfd_10g_bnkps", conf=cfg.face_det_conf, arcface_model="arcface_r100")
    eng = face.engine

    frames_np, ref_img = make_pool(args.pool, distractor_prob=args.distractor_prob)
    bank = PS.build_reference_bank(face, [ref_img], cfg)
    if bank is None:
        raise RuntimeError("reference image produced no face: cannot benchmark the matching stage")
    pool_dev = eng.to_device(frames_np)
    pool_pin = torch.from_numpy(frames_np).pin_memory()
    eng.sync()
    total = args.frames_per_step * world
    clip_dev = PooledDeviceClip(pool_dev, total)
    clip_host = PooledHostClip(pool_pin, total)

    faces_seen = {"n": 0, "passes": 0}

    def step(clip):
        stats = {}
        spans, bank_out = PS.prescan_batched(clip, 24, face, bank, cfg, batch=args.batch, stats=stats)
        faces_seen["digest"] = spans_digest(spans, bank_out)
        faces_seen["n"] = stats.get("faces", 0)               # faces aligned + embedded by THIS rank in the step
        faces_seen["passes"] = stats.get("arcface_passes", 0)  # ArcFace image passes (e(x), plus e(flip x) where the span logic needs it)
        faces_seen["spans"] = [list(map(int, sp)) for sp in spans]
        faces_seen["phase_ms"] = stats.get("phase_ms")
        faces_seen["bank"] = {k: stats.get(k) for k in ("bank_rows", "bank_versions", "distance_refreshes", "flip_on_demand", "early_flip_rows", "replay_parts_ms")}
        return spans

    def timed(clip, steps, profile=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.reset_launch_count()
        if profile:
            eng.profile_read(reset=True)
            eng.set_profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(eng.stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            spans = step(clip)
        e1.record(eng.stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        if world > 1:
            t = torch.tensor([ms], device=eng.tdev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        prof = None
        if profile:
            eng.set_profile(False)
            prof = eng.profile_read(reset=True)
        return ms, wall, spans, eng.launch_count(), prof

    for _ in range(args.warmup):
        step(clip_dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # headline region: K steps as a user runs them.  No per-launch CUDA events here: an event between two convolution launches
    # serialises them, which switches off the kernels' programmatic dependent launch (the prologue of launch i+1 under the tail of
    # launch i) and costs ~5 % by itself (gpurun r2u: 6 870 frames/s without events, 6 250 with them, same box, alternating)
    ms, wall, spans, launches, _ = timed(clip_dev, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    phases_dev = faces_seen.get("phase_ms")
    value = total * args.steps / (ms / 1000.0)
    # roofline region: the same K steps again with one CUDA-event pair around every convolution launch (on the engine's stream)
    ms_ev, _, _, _, prof = timed(clip_dev, args.steps, profile=True)
    value_ev = total * args.steps / (ms_ev / 1000.0)
    # faces through ArcFace per step (superset: every face is embedded with and without flip)
    face_passes = None
    # end to end: host frames, H2D inside the timed region
    for _ in range(2):            # the host-resident path has its own buffers (second ArcFace context, pinned staging) to warm up
        step(clip_host)
    clip_host.h2d_bytes = 0
    ms_e, _, _, _, _ = timed(clip_host, max(1, args.steps // 2))
    e2e_steps = max(1, args.steps // 2)
    e2e_value = total * e2e_steps / (ms_e / 1000.0)
    h2d = clip_host.h2d_bytes // e2e_steps

    # every rank must have replayed to the same spans / bank; rank 0 additionally repeats the whole clip ALONE (no collective)
    # and the N-rank result must equal it
    digest = faces_seen.get("digest")
    multi = None
    if world > 1:
        mine = torch.tensor([int(digest, 16) & 0x7fffffffffffffff], dtype=torch.int64, device=eng.tdev)
        allv = torch.empty((world,), dtype=torch.int64, device=eng.tdev)
        dist.all_gather_into_tensor(allv, mine)
        ranks_agree = bool((allv == allv[0]).all().item())
        single = None
        if rank == 0:
            s1, b1 = PS.prescan_batched(clip_dev, 24, face, bank, cfg, batch=args.batch, single_rank=True)
            single = spans_digest(s1, b1)
        multi = {"ranks_agree": ranks_agree, "single_rank_digest": single, "equals_single_rank": single == digest if rank == 0 else None}
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as fh:
            traffic = json.load(fh)
    except Exception:
        pass
    conv_ms, conv_flops, conv_n = prof
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    g_s = graphs.build_graph("scrfd_10g_bnkps")
    g_a = graphs.build_graph("arcface_r100")
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": args.frames_per_step, "batch": args.batch,
                   "pool_frames": args.pool, "distractor_prob": args.distractor_prob,
                   "l2_policy": "inputs larger than L2 (pool 398 MB of 1080p frames, cycled)",
                   "detector_input": 512, "kept_spans": faces_seen.get("spans"), "spans_sha": digest, "faces_per_step_per_gpu": faces_seen["n"], "arcface_passes_per_step_per_gpu": faces_seen["passes"],
                   "flip_tta": "e(flip x) only for faces evaluated while a span is active (as the reference); N>1 ranks embed both variants before the all-gather",
                   "weights": "SCRFD trained on synthetic faces; ArcFace seeded random + calibrated affine (no checkpoints offline)",
                   "scrfd_gflop_per_frame": 2e-9 * graphs.graph_macs(g_s, 256, 256),
                   "arcface_gflop_per_face_pass": 2e-9 * graphs.graph_macs(g_a, 112, 112)},
        "faces_embedded_per_sec": faces_seen["n"] * world * args.steps / (ms / 1000.0),
        "arcface_image_passes_per_sec": faces_seen["passes"] * world * args.steps / (ms / 1000.0),
        "gpu_launches": launches,
        "phase_ms_last_step": phases_dev, "phase_ms_last_e2e_step": faces_seen.get("phase_ms"), "bank_last_step": faces_seen.get("bank"),
        "wall_ms_per_step": 1000.0 * wall / args.steps,
        "value_with_launch_events": {"value": value_ev, "ms_per_step": ms_ev / args.steps,
                                     "note": "the roofline region: same K steps with a CUDA-event pair around every convolution launch"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
                # device->host per step, from the sizes of the tensors the step reads back: per face its box / quality / count
                # (16 + 8 + 4 B), its distance to the initial bank for the flip prediction (4 B) and the final plain + flip
                # distances for the refine probes (8 B); per frame batch the three detection counts; per distance refresh the
                # plain + flip similarities of the rows still ahead (upper bound: all rows); per accepted bank offer one
                # 2 KB feature row (the features themselves stay on the device)
                "d2h_bytes_per_step": int(faces_seen["n"] * (16 + 8 + 4 + 4 + 8) + (args.frames_per_step // args.batch + 1) * 3 * 4 * args.batch
                                          + (faces_seen.get("bank") or {}).get("distance_refreshes", 0) * faces_seen["n"] * 8
                                          + ((faces_seen.get("bank") or {}).get("bank_versions", 0) or 0) * 2048)},
        "roofline": {"kernel": "conv_tc2_kernel", "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf if peak_tf else None,
                     # DRAM bytes per launch of the dominant layer shape, read from the committed ncu --set full capture
                     # (dram__bytes_read.sum + dram__bytes_write.sum; profiles/r02_ncu_traffic.json names the command)
                     "traffic": traffic.get("traffic_bytes_per_launch"), "traffic_layer": traffic.get("layer"),
                     "traffic_source": "profiles/r02_ncu_traffic.json" if traffic else None,
                     "peak_source": peak_src,
                     "measured_in": "second timed region of the same K steps, CUDA-event pair per launch (value_with_launch_events)",
                     "conv_launches": conv_n, "conv_ms_per_step": conv_ms / args.steps,
                     "conv_share_of_step": (conv_ms / ms_ev) if ms_ev > 0 else None},
    }
    if multi is not None:
        line["multi_gpu_check"] = multi
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ora = {}
        rate, dt = cpu_oracle_rate(frames_np, ref_img, cfg, args.cpu_sample, threads, out=ora)
        line["cpu_baseline"] = {"value": rate, "unit": "frames/s", "cores": threads, "kind": "port",
                                "sample": f"first {args.cpu_sample} frames of the same clip, CPU oracle "
                                          f"(torch-CPU fp32 + cv2), {dt:.1f} s after a 4-frame warm-up"}
        # parity on the bench configuration itself: the GPU path over the SAME sample clip vs the oracle run just timed
        ns = args.cpu_sample
        glog = []
        gspans, gbank = PS.prescan_batched(PooledDeviceClip(pool_dev, ns), 24, face, bank, cfg, batch=args.batch, log=glog, single_rank=True)
        olog = ora["log"]
        same = len(glog) == len(olog) and all(g["idx"] == o["idx"] and g["skip"] == o["skip"] and g["nfaces"] == o["nfaces"]
                                              for g, o in zip(glog, olog))
        dfd = [abs(g["best"] - o["best"]) for g, o in zip(glog, olog) if g["nfaces"] == o["nfaces"] and o["nfaces"] > 0]
        cosb = [float(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))) for a, b in zip(np.asarray(bank), np.asarray(ora["bank0"]))]
        line["parity"] = {"oracle": "CPU restatement (torch fp32 + cv2), same weights, same frames", "samples": ns,
                          "decisions_equal": bool(same), "max_abs_dfd": float(max(dfd)) if dfd else None,
                          "spans_equal": [list(map(int, sp)) for sp in gspans] == [list(map(int, sp)) for sp in ora["spans"]],
                          "bank_rows_equal": int(np.asarray(gbank).shape[0]) == int(np.asarray(ora["bank"]).shape[0]),
                          "ref_bank_min_cos": min(cosb) if cosb else None, "tolerance_dfd": 3e-3}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
