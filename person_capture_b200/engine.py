"""Device engine: one libpcb200 context per GPU, graphs loaded, torch tensors as device buffers.

PyTorch is used only for device memory and streams (torch.Tensor.data_ptr() crosses the C ABI,
as the reference already does for its ORT io_binding, face_embedder.py:1321-1340); every kernel
on the path is in libpcb200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from . import _lib as L
from . import graphs, weights


@dataclass
class DetectResult:
    """Per-pass detector outputs (device tensors; call .cpu() views after Engine.sync())."""
    det: torch.Tensor        # [n, max_det, 5]
    kps: torch.Tensor        # [n, max_det, 10]
    raw_count: torch.Tensor  # [n]
    acc_box: torch.Tensor    # [n, max_det, 4] int32
    acc_kps: torch.Tensor    # [n, max_det, 10]
    acc_score: torch.Tensor  # [n, max_det]
    acc_count: torch.Tensor  # [n]
    acc_unfiltered: torch.Tensor  # [n] entries before the min-size filter
    counts: Optional[torch.Tensor] = None   # [3, n] = raw_count | acc_count | acc_unfiltered when allocated by Engine.alloc_detect


@dataclass
class AlignResult:
    face_count: torch.Tensor   # [n]
    face_total: torch.Tensor   # [1]
    face_frame: torch.Tensor   # [max_faces]
    face_box: torch.Tensor     # [max_faces, 4]
    face_kind: torch.Tensor    # [max_faces]
    chips: torch.Tensor        # [max_faces, 112, 112, 3] uint8
    quality: torch.Tensor      # [max_faces] float64


class Engine:
    def __init__(self, device: int = 0, scrfd: Optional[str] = "scrfd_10g_bnkps", arcface: Optional[str] = "arcface_r100",
                 scrfd_params=None, arcface_params=None, lib_path: Optional[str] = None, stream_priority: int = 0):
        if not torch.cuda.is_available():
            raise L.PcbError("no CUDA device: the identity path has no CPU fallback")
        self.lib = L.load(lib_path)
        self.device = int(device)
        self.tdev = torch.device("cuda", self.device)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.Stream(device=self.tdev, priority=int(stream_priority))
        self.copy_stream = torch.cuda.Stream(device=self.tdev)   # H2D prefetch of the next frame batch (prescan.compute_superset)
        self.ctx = self.lib.pcb_create(self.device, C.c_void_p(self.stream.cuda_stream))
        if not self.ctx:
            raise L.PcbError("pcb_create failed (needs an sm_100 device)")
        self.graphs = {}
        self.bank_rows = 0
        self.scrfd_name, self.arcface_name = scrfd, arcface
        if scrfd:
            self.load_model(L.MODEL_SCRFD, scrfd, scrfd_params)
        if arcface:
            self.load_model(L.MODEL_ARCFACE, arcface, arcface_params)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.pcb_last_error(self.ctx)
            raise L.PcbError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.pcb_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.lib.pcb_sync(self.ctx), "pcb_sync")

    def set_conv_impl(self, impl: int):
        self._check(self.lib.pcb_set_conv_impl(self.ctx, int(impl)), "pcb_set_conv_impl")

    def launch_count(self) -> int:
        return int(self.lib.pcb_launch_count(self.ctx))

    def reset_launch_count(self):
        self.lib.pcb_reset_launch_count(self.ctx)

    def set_profile(self, on: bool):
        self._check(self.lib.pcb_set_profile(self.ctx, 1 if on else 0), "pcb_set_profile")

    def profile_read(self, reset: bool = True):
        """-> (conv kernel ms, algorithmic conv FLOPs, conv launches) since the last reset."""
        ms, fl, n = C.c_double(), C.c_double(), C.c_longlong()
        self._check(self.lib.pcb_profile_read(self.ctx, C.byref(ms), C.byref(fl), C.byref(n), 1 if reset else 0), "pcb_profile_read")
        return ms.value, fl.value, n.value

    def empty(self, shape, dtype):
        # allocate under the engine's stream so the caching allocator orders reuse against our kernels
        with torch.cuda.stream(self.stream):
            return torch.empty(shape, dtype=dtype, device=self.tdev)

    def zeros(self, shape, dtype):
        with torch.cuda.stream(self.stream):
            return torch.zeros(shape, dtype=dtype, device=self.tdev)

    def to_device(self, arr: np.ndarray, stream=None) -> torch.Tensor:
        """Host -> device on the engine's stream (or `stream`) through pinned staging."""
        t = torch.from_numpy(np.ascontiguousarray(arr))
        with torch.cuda.stream(stream if stream is not None else self.stream):
            return t.pin_memory().to(self.tdev, non_blocking=True)

    # ------------------------------------------------------------------ graphs
    def load_model(self, slot: int, name: str, params=None):
        g = graphs.build_graph(name)
        P = params if params is not None else weights.load_params(name)
        ops, blob, outs, reg = graphs.pack(g, P)
        buf = C.create_string_buffer(blob, len(blob))
        regp = reg.ctypes.data_as(C.POINTER(C.c_float)) if reg is not None else None
        self._check(self.lib.pcb_model_load(self.ctx, slot, ops, len(g.ops), g.n_tensors, C.cast(buf, C.c_void_p), len(blob),
                                            outs, len(g.outputs), regp), f"pcb_model_load({name})")
        self.graphs[slot] = g

    def get_tensor(self, slot: int, tid: int) -> np.ndarray:
        n, c, h, w = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.pcb_model_get_tensor(self.ctx, slot, tid, None, C.byref(n), C.byref(c), C.byref(h), C.byref(w)),
                    "pcb_model_get_tensor")
        out = np.empty((n.value, c.value, h.value, w.value), np.float32)
        self._check(self.lib.pcb_model_get_tensor(self.ctx, slot, tid, out.ctypes.data_as(C.c_void_p), C.byref(n), C.byref(c),
                                                  C.byref(h), C.byref(w)), "pcb_model_get_tensor")
        return out

    # ------------------------------------------------------------------ K0
    def resize(self, frames: torch.Tensor, nh: int, nw: int, area: bool = True) -> torch.Tensor:
        n, h, w, _ = frames.shape
        out = self.empty((n, nh, nw, 3), torch.uint8)
        fn = self.lib.pcb_resize_area if area else self.lib.pcb_resize_linear
        self._check(fn(self.ctx, frames.data_ptr(), n, h, w, out.data_ptr(), nh, nw), "pcb_resize")
        return out

    # ------------------------------------------------------------------ K1+K2+K3
    def alloc_detect(self, n: int, max_det: int) -> DetectResult:
        e = self.empty
        counts = e((3, n), torch.int32)     # raw / accumulated / unfiltered counts share one buffer: one device->host copy reads all
        res = DetectResult(e((n, max_det, 5), torch.float32), e((n, max_det, 10), torch.float32), counts[0],
                           e((n, max_det, 4), torch.int32), e((n, max_det, 10), torch.float32), e((n, max_det), torch.float32),
                           counts[1], counts[2])
        res.counts = counts
        return res

    def _detect_args(self, frames, S, thresh, rot, pad, fix_mode, fix_scale_inv, orig_hw, min_box, res: DetectResult):
        n, h, w, _ = frames.shape
        a = L.DetectArgs()
        a.frames_dev = frames.data_ptr()
        a.n, a.h, a.w, a.S = n, h, w, int(S)
        a.det_thresh = float(thresh)
        a.rot_deg, a.pad_replicate, a.fix_mode, a.fix_scale_inv = int(rot), int(pad), int(fix_mode), float(fix_scale_inv)
        a.orig_h, a.orig_w = (int(orig_hw[0]), int(orig_hw[1])) if orig_hw else (h, w)
        a.min_box_px = int(min_box)
        a.max_det = res.det.shape[1]
        a.det_dev, a.kps_dev, a.raw_count_dev = res.det.data_ptr(), res.kps.data_ptr(), res.raw_count.data_ptr()
        a.acc_box_dev, a.acc_kps_dev = res.acc_box.data_ptr(), res.acc_kps.data_ptr()
        a.acc_score_dev, a.acc_count_dev = res.acc_score.data_ptr(), res.acc_count.data_ptr()
        a.acc_unfiltered_dev = res.acc_unfiltered.data_ptr()
        return a

    def detect(self, frames: torch.Tensor, S: int, thresh: float, rot: int = 0, pad: int = 0, fix_mode: int = L.FIX_NONE,
               fix_scale_inv: float = 1.0, orig_hw=None, min_box: int = 8, max_det: int = 256,
               out: Optional[DetectResult] = None) -> DetectResult:
        """One SCRFD pass over frames uint8 [n,h,w,3] (device)."""
        assert frames.is_cuda and frames.dtype == torch.uint8 and frames.is_contiguous()
        res = out if out is not None else self.alloc_detect(frames.shape[0], max_det)
        a = self._detect_args(frames, S, thresh, rot, pad, fix_mode, fix_scale_inv, orig_hw, min_box, res)
        self._check(self.lib.pcb_detect(self.ctx, C.byref(a)), "pcb_detect")
        return res

    def letterbox(self, frames: torch.Tensor, S: int, rot: int = 0, pad: int = 0, want_det_img: bool = False):
        n, h, w, _ = frames.shape
        half = S // 2
        out = self.zeros((n, half + L.P_PAD, half + L.P_PAD, 32), torch.float16)
        det = self.empty((n, S, S, 3), torch.uint8) if want_det_img else None
        self._check(self.lib.pcb_letterbox(self.ctx, frames.data_ptr(), n, h, w, S, rot, pad, out.data_ptr(),
                                           det.data_ptr() if det is not None else None), "pcb_letterbox")
        return out, det

    def decode_nms(self, heads: List[torch.Tensor], reg_scale, S: int, thresh: float, det_scale: float, frame_hw,
                   rot=0, pad=0, fix_mode=L.FIX_NONE, fix_scale_inv=1.0, min_box=8, max_det=256) -> DetectResult:
        n = heads[0].shape[0]
        res = self.alloc_detect(n, max_det)
        a = L.DetectArgs()
        a.n, a.h, a.w, a.S = n, int(frame_hw[0]), int(frame_hw[1]), int(S)
        a.det_thresh = float(thresh)
        a.rot_deg, a.pad_replicate, a.fix_mode, a.fix_scale_inv = int(rot), int(pad), int(fix_mode), float(fix_scale_inv)
        a.orig_h, a.orig_w = int(frame_hw[0]), int(frame_hw[1])
        a.min_box_px, a.max_det = int(min_box), max_det
        a.det_dev, a.kps_dev, a.raw_count_dev = res.det.data_ptr(), res.kps.data_ptr(), res.raw_count.data_ptr()
        a.acc_box_dev, a.acc_kps_dev = res.acc_box.data_ptr(), res.acc_kps.data_ptr()
        a.acc_score_dev, a.acc_count_dev = res.acc_score.data_ptr(), res.acc_count.data_ptr()
        a.acc_unfiltered_dev = res.acc_unfiltered.data_ptr()
        rs = np.asarray(reg_scale, np.float32)
        torch.cuda.current_stream().synchronize()   # heads may have been produced on torch's stream
        self._check(self.lib.pcb_decode_nms(self.ctx, heads[0].data_ptr(), heads[1].data_ptr(), heads[2].data_ptr(),
                                            rs.ctypes.data_as(C.POINTER(C.c_float)), C.byref(a), float(det_scale)),
                    "pcb_decode_nms")
        return res

    # ------------------------------------------------------------------ K4
    def align(self, frames: torch.Tensor, det: DetectResult, max_faces: int) -> AlignResult:
        n, h, w, _ = frames.shape
        e = self.empty
        res = AlignResult(e((n,), torch.int32), e((1,), torch.int32), e((max_faces,), torch.int32),
                          e((max_faces, 4), torch.int32), e((max_faces,), torch.int32),
                          e((max_faces, L.CHIP, L.CHIP, 3), torch.uint8), e((max_faces,), torch.float64))
        a = L.AlignArgs()
        a.frames_dev = frames.data_ptr()
        a.n, a.h, a.w, a.max_det = n, h, w, det.acc_box.shape[1]
        a.acc_box_dev, a.acc_kps_dev = det.acc_box.data_ptr(), det.acc_kps.data_ptr()
        a.acc_score_dev, a.acc_count_dev = det.acc_score.data_ptr(), det.acc_count.data_ptr()
        a.max_faces = max_faces
        a.face_count_dev, a.face_total_dev = res.face_count.data_ptr(), res.face_total.data_ptr()
        a.face_frame_dev, a.face_box_dev, a.face_kind_dev = res.face_frame.data_ptr(), res.face_box.data_ptr(), res.face_kind.data_ptr()
        a.chips_dev, a.quality_dev = res.chips.data_ptr(), res.quality.data_ptr()
        self._check(self.lib.pcb_align(self.ctx, C.byref(a)), "pcb_align")
        return res

    # ------------------------------------------------------------------ ArcFace + K5
    def embed(self, chips: torch.Tensor, f: int, flip):
        """flip False: e(x); True: e(x) and e(flip x); "only": e(flip x) alone (returned as the second element)."""
        only = flip == "only"
        emb = None if only else self.empty((max(f, 1), L.FEAT_DIM), torch.float32)
        emb_flip = self.empty((max(f, 1), L.FEAT_DIM), torch.float32) if flip else None
        if f > 0:
            self._check(self.lib.pcb_embed(self.ctx, chips.data_ptr(), f, emb.data_ptr() if emb is not None else None,
                                           emb_flip.data_ptr() if emb_flip is not None else None), "pcb_embed")
        return emb, emb_flip

    def set_bank(self, bank: Optional[np.ndarray], token=None):
        """Upload the reference bank.  `token` (any hashable that is unique to the bank CONTENT, e.g. (RefBank.serial, version) -- never id(obj): addresses are reused): skip the upload when the
        bank on the device already carries it -- a 10 000-row bank is 20 MB."""
        if token is not None and token == getattr(self, "_bank_token", None):
            return
        self._bank_token = None          # a failed upload must not look like a cached one
        if bank is None or np.asarray(bank).size == 0:
            self._check(self.lib.pcb_set_bank(self.ctx, None, 0), "pcb_set_bank")
            self.bank_rows = 0
        else:
            b = np.ascontiguousarray(np.asarray(bank, np.float32).reshape(-1, L.FEAT_DIM))
            self._check(self.lib.pcb_set_bank(self.ctx, b.ctypes.data_as(C.c_void_p), b.shape[0]), "pcb_set_bank")
            self.bank_rows = b.shape[0]
        self._bank_token = token

    def match(self, emb: torch.Tensor, emb_flip: Optional[torch.Tensor], use_flip: Optional[torch.Tensor], f: int,
              want_feat: bool = True):
        """-> (feat [f,512] or None, sim [f], argmax [f]) device tensors; fd = 1 - sim."""
        feat = self.empty((max(f, 1), L.FEAT_DIM), torch.float32) if want_feat else None
        sim = self.empty((max(f, 1),), torch.float32)
        arg = self.empty((max(f, 1),), torch.int32)
        if f > 0:
            self._check(self.lib.pcb_match(self.ctx, emb.data_ptr(), emb_flip.data_ptr() if emb_flip is not None else None,
                                           use_flip.data_ptr() if use_flip is not None else None, f,
                                           feat.data_ptr() if feat is not None else None, sim.data_ptr(), arg.data_ptr()), "pcb_match")
        return feat, sim, arg
