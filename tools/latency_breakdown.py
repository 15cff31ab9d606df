"""Where a lock-face ROI frame's time goes (config 5): CUDA-event time of each stage of one extract(roi, imgsz=1280) + distance,
and the wall time of the whole step.  usage: python tools/latency_breakdown.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from person_capture_b200 import synth, prescan as PS, mainpass as MP, _lib as L
from person_capture_b200.face_embedder import FaceEmbedder
from person_capture_b200.params import PrescanParams

face = FaceEmbedder("cuda:0", "scrfd_10g_bnkps", conf=0.5, arcface_model="arcface_r100")
eng = face.engine
clip = synth.ClipSpec(3840, 2160, 8, seed=1005, target=1, others=(), target_segments=[(0, 7)], face_px=(60, 110))
fr, truth = clip.frame_with_truth(3)
x1, y1, x2, y2 = [float(v) for v in truth[0][1]]
rx1, ry1, rx2, ry2 = MP.expand_xyxy((x1, y1, x2, y2), max(16.0, (x2 - x1) * 1.25), max(16.0, (y2 - y1) * 1.25), 3840, 2160)
roi = eng.to_device(np.ascontiguousarray(fr[ry1:ry2, rx1:rx2]))[None].contiguous()
print("roi", roi.shape)
bank = np.random.default_rng(0).normal(size=(64, 512)).astype(np.float32)
eng.set_bank(bank)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record(eng.stream)
    return e


for it in range(6):
    t0 = time.perf_counter()
    e = [ev()]
    det = eng.detect(roi, 1280, 0.5, max_det=256); e.append(ev())
    al = eng.align(roi, det, max_faces=256); e.append(ev())
    emb, embf = eng.embed(al.chips, 1, True); e.append(ev())
    feat, sim, arg = eng.match(emb, embf, None, 1); e.append(ev())
    t1 = time.perf_counter()
    eng.sync()
    t2 = time.perf_counter()
    ms = [e[i].elapsed_time(e[i + 1]) for i in range(4)]
    print(f"it {it}: gpu detect {ms[0]:.3f} align {ms[1]:.3f} embed {ms[2]:.3f} match {ms[3]:.3f} | host enqueue {1e3*(t1-t0):.3f} ms, total wall {1e3*(t2-t0):.3f} ms, launches {eng.launch_count()}")
    eng.reset_launch_count()
# through the public API
for it in range(5):
    eng.sync(); t0 = time.perf_counter()
    faces = face.extract(roi[0], imgsz=1280)
    t1 = time.perf_counter()
    print(f"extract: {1e3*(t1-t0):.3f} ms, faces {len(faces)}")
