"""FaceEmbedder -- drop-in for the reference class on the SCRFD + ArcFace configuration.

Mirrors the public surface of person_capture/face_embedder.py::FaceEmbedder (ctor :383-387,
extract :1663, best_face :2504, set_prescan_fast :1224, set_prescan_hint :1233,
configure_rotation_strategy :1238, and the attributes the pre-scan mutates, gui_app.py:1162-1183),
but every detector / alignment / embedding step runs in libpcb200 on the GPU.  The host keeps
only the reference's per-call *policy* (which SCRFD passes to run, face_embedder.py:2189-2208,
2251-2433) because it is a handful of scalar decisions that depend on per-pass detection
counts.  There is no CPU fallback: without the CUDA library the constructor raises.

Out of scope: the YOLOv8-face and OpenCLIP branches, downloaders.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from . import _lib as L
from .engine import Engine

_ROT_PAD = 24          # face_embedder.py:2393
_MIN_SIDE = 320        # smallest SCRFD square (face_embedder.py:2204)


def round32(x: int) -> int:
    return ((int(x) + 31) // 32) * 32


@dataclass
class _Pass:
    """One SCRFD invocation of the reference's per-frame schedule."""
    frames: torch.Tensor
    size: int
    conf: float
    rot: int = 0
    pad: int = 0
    fix: int = L.FIX_NONE
    scale_inv: float = 1.0


class FaceEmbedder:
    def __init__(self, ctx: str = "cuda", yolo_model: str = "scrfd_10g_bnkps", conf: float = 0.30, use_arcface: bool = True,
                 clip_model_name: str = "ViT-L-14", clip_pretrained: str = "laion2b_s32b_b82k", progress=None,
                 trt_lib_dir: Optional[str] = None, *, arcface_model: str = "arcface_r100", engine: Optional[Engine] = None,
                 rot_phase: int = 0, max_det: int = 256):
        name = str(yolo_model)
        if not name.lower().startswith("scrfd"):
            raise RuntimeError("person_capture_b200 implements the SCRFD detector backend only "
                               "(yolov8-face is outside the hot path, SURVEY.md section 2)")
        if not use_arcface:
            raise RuntimeError("person_capture_b200 implements the ArcFace embedding backend only")
        for suffix in (".onnx",):
            if name.endswith(suffix):
                name = name[: -len(suffix)]
        dev = 0
        s = str(ctx).lower()
        if s.startswith("cuda:"):
            dev = int(s.split(":", 1)[1])
        elif not s.startswith("cuda"):
            raise RuntimeError("SCRFD backend requires a CUDA ctx (as the reference, face_embedder.py:510-515)")
        self.engine = engine if engine is not None else Engine(dev, scrfd=name, arcface=arcface_model)
        self.backend = "scrfd"
        self.detector_backend = "scrfd"
        self.use_arcface = True
        self.conf = float(conf)
        self.progress = progress
        self.max_det = int(max_det)
        # knobs and state: same names and defaults as the reference (face_embedder.py:473-500)
        self.scrfd_tta_scales = (0.75, 0.60)
        self.scrfd_probe_conf_cap = 0.20
        self.scrfd_edge_pad_frac = 0.06
        self.scrfd_min_box_px = 8
        self._fast_prescan = False
        self._prescan_rr = 0
        self._prescan_rr_mode = "rr"
        self._prescan_escalate = False
        self._probe_conf = 0.03
        self._high_90 = 1536
        self._high_180 = 1280
        self._prescan_period = 3
        self._prescan_probe_imgsz = 384
        self._prescan_no_upscale_det = True
        self._heavy_cap = 2048
        self._frame_idx = 0
        self._no_face_streak = 0
        self._last_face_idx = -10 ** 9
        self._rot_cycle = 0
        self.rot_adaptive = True
        self.rot_every_n = 12
        self.rot_after_hit_frames = 8
        self.fast_no_face_imgsz = 512
        self.rot_phase = int(rot_phase) & 7   # stands in for `id(self) & 7` (face_embedder.py:2338)
        self.keep_debug = False               # True: extract() also keeps last_chips / last_kinds (device->host copies)
        self.last_passes: List[dict] = []

    def aux_engine(self) -> Engine:
        """A second context on the same GPU with the ArcFace graph only (own stream, own activation set).  The host-resident
        pre-scan embeds on it, so that ArcFace runs are not queued behind SCRFD passes that wait for their frames to cross
        PCIe (prescan.EarlyFlips)."""
        if getattr(self, "_aux_engine", None) is None:
            self._aux_engine = Engine(self.engine.device, scrfd=None, arcface=self.engine.arcface_name)
        return self._aux_engine

    # ---- knobs -----------------------------------------------------------------------
    def set_prescan_fast(self, enable: bool, *, mode: str = "rr") -> None:
        self._fast_prescan = bool(enable)
        self._prescan_rr_mode = str(mode)
        if enable:
            self._prescan_rr = 0

    def set_prescan_hint(self, *, escalate: bool = False) -> None:
        self._prescan_escalate = bool(escalate)

    def configure_rotation_strategy(self, *, adaptive=None, every_n=None, after_hit_frames=None, fast_no_face_imgsz=None) -> None:
        if adaptive is not None:
            self.rot_adaptive = bool(adaptive)
        if every_n is not None:
            self.rot_every_n = max(1, int(every_n))
        if after_hit_frames is not None:
            self.rot_after_hit_frames = max(0, int(after_hit_frames))
        if fast_no_face_imgsz is not None:
            self.fast_no_face_imgsz = max(0, int(fast_no_face_imgsz))
        self._rot_cycle = 0

    @staticmethod
    def best_face(faces):
        if not faces:
            return None
        return max(faces, key=lambda f: (f["quality"], (f["bbox"][2] - f["bbox"][0]) * (f["bbox"][3] - f["bbox"][1])))

    # ---- size policy (face_embedder.py:2189-2208) --------------------------------------
    def upright_size(self, H0: int, W0: int, imgsz: Optional[int]) -> int:
        dyn = int(imgsz) if (imgsz is not None and imgsz > 0) else 640
        if self._no_face_streak >= 3:
            dyn = min(dyn, self.fast_no_face_imgsz)
        if self._fast_prescan:
            dyn = min(dyn, int(self._prescan_probe_imgsz))
            if self._prescan_no_upscale_det:
                dyn = min(dyn, max(_MIN_SIDE, (max(H0, W0) // 32) * 32))
        return round32(max(_MIN_SIDE, dyn))

    def heavy_sizes(self, H0: int, W0: int, dyn: int):
        L_ = max(H0, W0)
        cap = max(int(self._heavy_cap), dyn)
        return (min(round32(max(dyn, int(0.75 * L_))), cap), min(round32(max(dyn, int(0.67 * L_))), cap))

    def probe_size(self, dyn: int) -> int:
        return round32(max(_MIN_SIDE, min(dyn, int(self._prescan_probe_imgsz))))

    def probe_conf_value(self) -> float:
        return max(0.02, float(self._probe_conf))

    def rotated_conf(self, deg: int) -> float:
        return max(0.10, float(self.conf) * (0.8 if deg in (90, 270) else 0.6))

    def heavy_size_fast(self, deg: int, heavy90: int, heavy180: int) -> int:
        heavy = heavy180 if deg == 180 else heavy90
        override = self._high_180 if deg == 180 else self._high_90
        if override and override > 0:
            heavy = max(heavy, round32(int(override)))
        return min(heavy, int(self._heavy_cap))

    # ---- one pass ----------------------------------------------------------------------
    def _run(self, ps: _Pass, H0: int, W0: int, min_box: Optional[int] = None):
        eng = self.engine
        mb = int(self.scrfd_min_box_px) if min_box is None else int(min_box)
        det = eng.detect(ps.frames, ps.size, ps.conf, rot=ps.rot, pad=ps.pad, fix_mode=ps.fix, fix_scale_inv=ps.scale_inv,
                         orig_hw=(H0, W0), min_box=mb, max_det=self.max_det)
        eng.sync()
        cnt = det.counts.cpu()              # one copy: raw / accumulated / unfiltered
        raw, acc, self._last_unfiltered = int(cnt[0, 0]), int(cnt[1, 0]), int(cnt[2, 0])
        self.last_passes.append(dict(shape=tuple(ps.frames.shape[1:3]), size=ps.size, conf=float(ps.conf), rot=ps.rot,
                                     pad=ps.pad, n=raw, acc=acc))
        return det, raw, acc

    # ---- extract -----------------------------------------------------------------------
    def extract(self, bgr_img, *, imgsz: Optional[int] = None):
        if bgr_img is None:
            return []
        on_device = isinstance(bgr_img, torch.Tensor)
        if (bgr_img.numel() if on_device else bgr_img.size) == 0:
            return []
        if bgr_img.ndim != 3 or bgr_img.shape[2] != 3 or bgr_img.dtype != (torch.uint8 if on_device else np.uint8):
            raise ValueError("extract expects a uint8 BGR image [H, W, 3]")
        self._frame_idx += 1
        self.last_passes = []
        eng = self.engine
        H0, W0 = int(bgr_img.shape[0]), int(bgr_img.shape[1])
        # frames already resident in HBM (north_star) are used in place; host arrays go through pinned staging
        if on_device:
            # the caller may have produced the tensor on torch's current stream; everything below runs on the engine's
            eng.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(eng.stream):
                frame = bgr_img.contiguous()[None]          # (a crop view is compacted here, on the stream that reads it)
        else:
            frame = eng.to_device(bgr_img[None])
        dyn = self.upright_size(H0, W0, imgsz)
        heavy90, heavy180 = self.heavy_sizes(H0, W0, dyn)
        fast = self._fast_prescan

        chosen, _, n_acc = self._run(_Pass(frame, dyn, self.conf), H0, W0)

        # `if not dets` in the reference looks at the accumulated list before the min-size filter
        if self._last_unfiltered == 0 and not fast:
            probe_conf = min(float(self.conf), float(self.scrfd_probe_conf_cap))
            scales = tuple(self.scrfd_tta_scales) + ((1.25,) if max(W0, H0) <= 1920 else ())
            for s in scales:
                if s == 1.0:
                    continue
                nh, nw = int(np.rint(H0 * s)), int(np.rint(W0 * s))
                scaled = eng.empty((1, nh, nw, 3), torch.uint8)
                eng._check(eng.lib.pcb_resize_factor(eng.ctx, frame.data_ptr(), 1, H0, W0, scaled.data_ptr(), float(s), float(s),
                                                     1 if s < 1.0 else 0), "pcb_resize_factor")
                dyn_s = round32(min(self._heavy_cap, max(_MIN_SIDE, int(dyn * s))))
                chosen, _, n_acc = self._run(_Pass(scaled, dyn_s, probe_conf, fix=L.FIX_SCALE, scale_inv=1.0 / s), H0, W0)
                # the reference breaks this loop on the list *before* its min-size filter (face_embedder.py:2282, 2317)
                if self._last_unfiltered > 0:
                    break
            if self._last_unfiltered == 0:
                pad = int(round(min(64, float(self.scrfd_edge_pad_frac) * max(W0, H0))))
                if pad > 0:
                    chosen, _, n_acc = self._run(_Pass(frame, dyn, probe_conf, pad=pad, fix=L.FIX_PADPROBE), H0, W0)

        if n_acc == 0:
            self._no_face_streak += 1
            if self.rot_adaptive:
                recent = (self._frame_idx - self._last_face_idx) <= self.rot_after_hit_frames
                if not recent and getattr(self, "rot_consults", None) is not None:
                    self.rot_consults.append(self._frame_idx)     # the absolute call counter decided (main_pass_sharded checks these)
                need_rot = recent or ((self._frame_idx + self.rot_phase) % self.rot_every_n) == 0
            else:
                need_rot = True
            if fast:
                period = max(1, int(self._prescan_period))
                need_rot = need_rot or self._prescan_escalate or (((self._frame_idx + self._prescan_rr) % period) == 0)
            if not need_rot:
                return []
            self._rot_cycle += 1
            if fast:
                if self._prescan_rr_mode == "rr":
                    rot_seq = ((90, 270)[self._prescan_rr % 2],)
                    self._prescan_rr += 1
                else:
                    rot_seq = (90, 270)
            else:
                rot_seq = (90, 270, 180)
            for deg in rot_seq:
                _, hits, _ = self._run(_Pass(frame, self.probe_size(dyn), self.probe_conf_value(), rot=deg), H0, W0)
                if fast and hits == 0:
                    continue
                if fast:
                    sizes = [self.heavy_size_fast(deg, heavy90, heavy180)]   # do_heavy is implied by hits > 0
                else:
                    sizes = []
                    for base in (max(dyn, 1280), max(dyn, 1536)):
                        b = round32(base)
                        if b not in sizes:
                            sizes.append(b)
                got = None
                for sz in sizes:
                    # rotated passes are accumulated after the min-size filter ran (face_embedder.py:2317 precedes 2363)
                    d, raw, acc = self._run(_Pass(frame, sz, self.rotated_conf(deg), rot=deg, pad=_ROT_PAD, fix=L.FIX_UNPAD), H0, W0,
                                            min_box=0)
                    if raw > 0:
                        got = (d, acc)
                        break
                if got is None:
                    continue
                if got[1] > 0:
                    chosen, n_acc = got
                    break
            if n_acc == 0:
                return []
        else:
            self._no_face_streak = 0
            self._last_face_idx = self._frame_idx
            self._rot_cycle = 0

        do_flip = (not fast) or self._prescan_escalate
        return self._faces_from(frame, chosen, do_flip, n_acc)

    def _faces_from(self, frame: torch.Tensor, det, do_flip: bool, n_acc: int):
        """K4 -> ArcFace -> normalise for the accumulated detections of the chosen pass.  The cross-pass suppression keeps
        between 1 and n_acc of them (the best one always survives), so the whole chain is enqueued for n_acc candidates
        without asking the GPU how many K4 kept -- rows past the kept count are computed and ignored -- and the host waits ONCE."""
        eng = self.engine
        al = eng.align(frame, det, max_faces=self.max_det)
        cand = min(int(n_acc), self.max_det)
        emb, emb_flip = eng.embed(al.chips, cand, do_flip)
        feat, _sim, _arg = eng.match(emb, emb_flip, None, cand)
        eng.sync()
        f = int(al.face_total.cpu()[0])
        if f == 0:
            return []
        self.last_face_count = f
        boxes = al.face_box[:f].cpu().numpy()
        qual = al.quality[:f].cpu().numpy()
        feats = feat[:f].cpu().numpy()
        if self.keep_debug:                  # chips / alignment kinds of the last call (parity tooling only: 37 KB per face)
            self.last_chips = al.chips[:f].cpu().numpy()
            self.last_kinds = al.face_kind[:f].cpu().numpy()
        out = [dict(bbox=boxes[i].astype(np.int32).copy(), feat=feats[i].astype(np.float32).copy(), quality=float(qual[i]))
               for i in range(f)]
        self.last_order = sorted(range(f), key=lambda i: (out[i]["quality"], (out[i]["bbox"][2] - out[i]["bbox"][0]) *
                                                           (out[i]["bbox"][3] - out[i]["bbox"][1])), reverse=True)
        # device-resident features in the returned order (prescan_sequential feeds them to pcb_match)
        with torch.cuda.stream(eng.stream):
            self.last_feats_dev = feat[:f].index_select(0, torch.as_tensor(self.last_order, device=feat.device)).contiguous()
        return [out[i] for i in self.last_order]
