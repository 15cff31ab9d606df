"""Worker of tests/test_gpu_multirank.py (launched by torch.distributed.run, one rank per GPU): the N-rank pre-scan of a clip
must give every rank the spans / bank / per-sample log that ONE rank computes alone."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    out_path = sys.argv[1]
    eager = len(sys.argv) > 2 and sys.argv[2] == "eager"
    if eager:
        os.environ["PCB_EAGER_FLIP"] = "1"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from person_capture_b200 import prescan as PS, synth
    from person_capture_b200.face_embedder import FaceEmbedder
    from person_capture_b200.params import PrescanParams
    cfg = PrescanParams(face_model="scrfd_2.5g_bnkps", prescan_stride=1, prescan_max_width=416, prescan_fd_enter=0.62,
                        prescan_fd_exit=0.72, prescan_fd_add=0.50, face_quality_min=40.0, prescan_min_segment_sec=0.5,
                        prescan_pad_sec=0.25, prescan_bridge_gap_sec=0.25, prescan_exit_cooldown_sec=0.25,
                        prescan_boundary_refine_sec=0.5, prescan_add_cooldown_samples=2)
    clip = synth.ClipSpec(640, 360, 240, seed=1003)
    frames = np.stack([clip.frame(i) for i in range(clip.n_frames)])
    face = FaceEmbedder(f"cuda:{local}", "scrfd_2.5g_bnkps", conf=cfg.face_det_conf, arcface_model="arcface_r50")
    bank = PS.build_reference_bank(face, [synth.reference_image(1, 512, seed=1003)], cfg)
    dev = PS.DeviceClip(face.engine.to_device(frames))
    log_n, log_1 = [], []
    stats = {}
    spans_n, bank_n = PS.prescan_batched(dev, 24, face, bank, cfg, batch=16, log=log_n, stats=stats)
    spans_1, bank_1 = PS.prescan_batched(dev, 24, face, bank, cfg, batch=16, log=log_1, single_rank=True)
    same_log = len(log_n) == len(log_1) and all(
        a["idx"] == b["idx"] and a["skip"] == b["skip"] and a["nfaces"] == b["nfaces"] and a["active_before"] == b["active_before"]
        and a["best"] == b["best"] for a, b in zip(log_n, log_1))
    res = dict(rank=rank, world=world, spans_n=[list(map(int, s)) for s in spans_n], spans_1=[list(map(int, s)) for s in spans_1],
               bank_equal=bool(np.asarray(bank_n).shape == np.asarray(bank_1).shape and np.array_equal(bank_n, bank_1)),
               bank_rows=int(np.asarray(bank_n).shape[0]), bank_rows_initial=int(np.asarray(bank).shape[0]), same_log=bool(same_log),
               refreshes=stats.get("distance_refreshes"), phase_ms=stats.get("phase_ms"))
    # main pass over kept spans, sharded by span blocks vs sequential (config 3's "sharded across N GPUs")
    from person_capture_b200 import mainpass as MP
    mcfg = PrescanParams(face_model="scrfd_2.5g_bnkps", face_thresh=0.62, face_quality_min=40.0, face_fullframe_imgsz=640,
                         frame_stride=2, face_fullframe_cadence=6, lock_face_roi_max_misses=3)
    mspans = [(4, 40), (50, 70), (84, 120), (130, 170), (180, 239)]
    mstats = {}
    hits_n = MP.main_pass_sharded(dev, 24.0, mspans, face, bank, mcfg, stats=mstats)
    face2 = FaceEmbedder(f"cuda:{local}", "scrfd_2.5g_bnkps", conf=cfg.face_det_conf, engine=face.engine)
    hits_1 = MP.main_pass(dev, 24.0, mspans, face2, bank, mcfg)
    key = lambda h: (h["idx"], h["site"], tuple(h["face_box"]), round(h["fd"], 6))
    res.update(main_hits=len(hits_1), main_equal=[key(h) for h in hits_n] == [key(h) for h in hits_1], main_rounds=mstats.get("rounds"),
               main_sites=sorted({h["site"] for h in hits_1}))
    gathered = [None] * world
    dist.all_gather_object(gathered, res)          # test harness only (not the product path)
    if rank == 0:
        with open(out_path, "w") as fh:
            json.dump(gathered, fh)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
