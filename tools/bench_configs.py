"""Secondary BASELINE.json configurations (parity-test cases, not the bench line): measured once per round for the record.
  C3  SCRFD-10G + ArcFace R100 full-frame main pass at 4K, face_fullframe_imgsz=1280, flip-TTA          (frames/s)
  C4  crowded scene: 1080p frames with ~64 faces each, bank of 10 000 embeddings                        (faces embedded/s)
  C5  lock-face ROI path: sequential per-frame latency with ArcFace R100 and bank distance, bank 64 / 10 000 (p50/p99 ms)
usage: python tools/bench_configs.py [--frames 96]"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=96)
ap.add_argument("--iters", type=int, default=300)
args = ap.parse_args()
import torch
from person_capture_b200 import synth, prescan as PS, mainpass as MP
from person_capture_b200.face_embedder import FaceEmbedder
from person_capture_b200.params import PrescanParams

face = FaceEmbedder("cuda:0", "scrfd_10g_bnkps", conf=0.5, arcface_model="arcface_r100")
eng = face.engine
cfg = PrescanParams(face_model="scrfd_10g_bnkps", face_thresh=0.62, face_quality_min=40.0, face_fullframe_imgsz=1280,
                    frame_stride=1, face_fullframe_cadence=12)
ref = synth.reference_image(1, 512, seed=1003)
bank = PS.build_reference_bank(face, [ref], cfg)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record(eng.stream)
    return e


# ---------------------------------------------------------------- C3
pool = 12
clip = synth.ClipSpec(3840, 2160, pool, seed=1003, target=1, others=(2, 3, 4), distractor_prob=1.0, target_segments=[(0, pool - 1)])
frames = eng.to_device(np.stack([clip.frame(i) for i in range(pool)]))


class Pooled:
    def __init__(self, fr, n):
        self.fr, self.total_frames = fr, n

    def device_batch(self, e, idxs, stream=None):
        with torch.cuda.stream(e.stream):
            return self.fr.index_select(0, torch.as_tensor([i % self.fr.shape[0] for i in idxs], device=self.fr.device))


src = Pooled(frames, args.frames)
idxs = list(range(args.frames))
MP.fullframe_identity(src, idxs, face, bank, cfg, batch=8)      # warm-up at full size (activation sets are allocated per capacity)
eng.sync()
e0 = ev()
recs = MP.fullframe_identity(src, idxs, face, bank, cfg, batch=8)
e1 = ev()
eng.sync()
ms = e0.elapsed_time(e1)
nf = sum(r["n_faces"] for r in recs)
print(f"C3 main pass 4K -> S=1280, flip-TTA: {args.frames} frames, {nf} faces, {ms:.1f} ms -> {args.frames / ms * 1e3:.1f} frames/s, "
      f"{2 * nf / ms * 1e3:.0f} ArcFace image passes/s, accepted {sum(r['accept'] for r in recs)}")
del frames, src

# ---------------------------------------------------------------- C4
pool = 8
clip4 = synth.ClipSpec(1920, 1080, pool, seed=1004, crowd=64, target_segments=[(0, pool - 1)])
fr4 = eng.to_device(np.stack([clip4.frame(i) for i in range(pool)]))
rng = np.random.default_rng(1004)
big = rng.normal(size=(10000, 512)).astype(np.float32)
big /= np.linalg.norm(big, axis=1, keepdims=True)
big[:len(bank)] = bank
cfg4 = PrescanParams(face_model="scrfd_10g_bnkps", prescan_stride=1, prescan_max_width=1920, prescan_bank_max=10000)
src4 = Pooled(fr4, 64)
with PS._PrescanFaceMode(face, cfg4):
    face._prescan_probe_imgsz = 1280          # crowded faces are small: detect at 1280^2
    for rep in range(2):
        eng.sync()
        e0 = ev()
        records, table = PS.compute_superset(src4, list(range(64)), face, cfg4, batch=16)
        eng.set_bank(big)
        _, sim, arg = eng.match(table.plain, None, None, table.count, want_feat=False)
        _, sim2, _ = eng.match(table.flip, None, None, table.count, want_feat=False)
        e1 = ev()
        eng.sync()
    ms = e0.elapsed_time(e1)
    print(f"C4 crowded 1080p (S=1280), bank 10 000: 64 frames, {table.count} faces ({table.count / 64:.1f}/frame), both flip variants, {ms:.1f} ms "
          f"-> {64 / ms * 1e3:.1f} frames/s, {table.count / ms * 1e3:.0f} faces/s, {2 * table.count / ms * 1e3:.0f} ArcFace image passes/s")
del fr4

# ---------------------------------------------------------------- C5
clip5 = synth.ClipSpec(3840, 2160, 8, seed=1005, target=1, others=(), target_segments=[(0, 7)], face_px=(60, 110))
fr5 = eng.to_device(np.stack([clip5.frame(i) for i in range(8)]))
for name, bk in (("bank 64", np.vstack([bank, big[len(bank):64]])), ("bank 10 000", big)):
    c5 = PrescanParams(face_model="scrfd_10g_bnkps", face_thresh=0.62, face_quality_min=40.0, face_fullframe_imgsz=1280, frame_stride=1,
                       face_fullframe_cadence=10 ** 9, prescan_bank_max=10000)
    st = MP.MainPassIdentity(face, bk, c5)
    st.bank.rows = [r.copy() for r in bk]
    st.bank.version += 1
    lat, sites = [], []
    for it in range(args.iters + 20):
        fr = fr5[it % 8]
        eng.sync()
        t0 = time.perf_counter()
        log = []
        st.step(it, fr, log)
        eng.sync()
        if it >= 20:
            lat.append((time.perf_counter() - t0) * 1e3)
            sites.append(log[0]["site"])
    lat = np.array(lat)
    print(f"C5 lock-face ROI path, 4K frames, {name}: p50 {np.percentile(lat, 50):.2f} ms, p99 {np.percentile(lat, 99):.2f} ms per frame "
          f"({sum(s == 'lock_roi' for s in sites)}/{len(sites)} frames answered by the ROI site)")
