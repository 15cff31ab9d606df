"""Session-level drop-ins (SURVEY.md 8b level 2): keep the reference's own FaceEmbedder and replace only the two device
calls it makes.

  ScrfdSession.detect(img, input_size=(S, S)) -> (det float32[F,5], kpss float32[F,5,2])
      replaces InsightFace `scrfd.detect(img_bgr_uint8, input_size=(S,S))` with `scrfd.det_thresh` set before each call
      (person_capture/face_embedder.py:2176-2187); boxes / landmarks in source-image pixels, NMS order, as upstream.
  ArcSession.run(None, {name: float32[n,3,112,112]}) -> [float32[n,512]]
      replaces `arc_sess.run` (face_embedder.py:1369) on pre-processed blobs ((RGB - 127.5) / 127.5, NCHW); also
      ArcSession.run_chips(uint8 BGR chips) which skips the host-side _arcface_preprocess.

Both ride on one Engine (one libpcb200 context); there is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .engine import Engine


class ScrfdSession:
    """`det_thresh` / `nms_thresh` attributes and `detect()` as insightface.model_zoo.scrfd.SCRFD exposes them."""

    def __init__(self, engine: Engine, max_det: int = 1024):
        if L.MODEL_SCRFD not in engine.graphs:
            raise L.PcbError("ScrfdSession needs an engine with a SCRFD graph loaded")
        self.engine = engine
        self.det_thresh = 0.5
        self.nms_thresh = 0.4          # fixed in the kernel (upstream default; the reference never changes it)
        self.max_det = int(max_det)

    def detect(self, img, input_size: Optional[Tuple[int, int]] = None, max_num: int = 0, metric: str = "default"):
        if input_size is None or input_size[0] != input_size[1] or input_size[0] % 32:
            raise ValueError("input_size must be a square multiple of 32 (the reference always passes (S, S), face_embedder.py:2185)")
        if max_num:
            raise ValueError("max_num > 0 (area/centre re-ranking) is not used by the reference and not implemented")
        eng = self.engine
        on_dev = isinstance(img, torch.Tensor)
        if on_dev:
            eng.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(eng.stream):
                frame = img.contiguous()[None]
        else:
            frame = eng.to_device(np.ascontiguousarray(img)[None])
        res = eng.detect(frame, int(input_size[0]), float(self.det_thresh), min_box=0, max_det=self.max_det)
        eng.sync()
        n = int(res.raw_count.cpu()[0])
        det = res.det[0, :n].cpu().numpy().astype(np.float32)
        kps = res.kps[0, :n].cpu().numpy().astype(np.float32).reshape(n, 5, 2)
        return det, kps


class _IOName:
    def __init__(self, name: str, shape):
        self.name, self.shape, self.type = name, shape, "tensor(float)"


class ArcSession:
    """The slice of onnxruntime.InferenceSession the reference touches on the ArcFace path: get_inputs / get_outputs / run."""

    def __init__(self, engine: Engine, input_name: str = "input.1"):
        if L.MODEL_ARCFACE not in engine.graphs:
            raise L.PcbError("ArcSession needs an engine with an ArcFace graph loaded")
        self.engine = engine
        self._in = _IOName(input_name, ["N", 3, L.CHIP, L.CHIP])
        self._out = _IOName("embedding", ["N", L.FEAT_DIM])

    def get_inputs(self):
        return [self._in]

    def get_outputs(self):
        return [self._out]

    def run(self, output_names, feeds: Dict[str, np.ndarray]) -> List[np.ndarray]:
        """feeds: {input name: float32 [n,3,112,112] = (RGB - 127.5) / 127.5}.  The blob is exactly what
        _arcface_preprocess emits from a uint8 chip, so the uint8 BGR chip is recovered without loss (x * 127.5 + 127.5 is
        an integer to 1e-5 for every uint8 input) and goes through the same device pre-processing as run_chips."""
        if len(feeds) != 1:
            raise ValueError("ArcSession.run expects exactly one input tensor")
        x = np.asarray(next(iter(feeds.values())), np.float32)
        if x.ndim != 4 or x.shape[1:] != (3, L.CHIP, L.CHIP):
            raise ValueError("ArcFace input must be float32 [n, 3, 112, 112]")
        px = x * np.float32(127.5) + np.float32(127.5)
        rgb = np.rint(px)
        if float(np.abs(px - rgb).max(initial=0.0)) > 1e-3 or rgb.min(initial=0.0) < 0 or rgb.max(initial=0.0) > 255:
            raise ValueError("ArcSession.run accepts blobs produced by _arcface_preprocess from uint8 chips only")
        chips = np.ascontiguousarray(rgb.astype(np.uint8).transpose(0, 2, 3, 1)[..., ::-1])      # NCHW RGB -> NHWC BGR
        return [self.run_chips(chips)]

    def run_chips(self, chips_bgr: np.ndarray, flip: bool = False):
        """uint8 BGR [n,112,112,3] -> raw embeddings float32 [n,512] (and e(flip x) when flip=True)."""
        eng = self.engine
        n = int(chips_bgr.shape[0])
        if n == 0:
            z = np.zeros((0, L.FEAT_DIM), np.float32)
            return (z, z) if flip else z
        emb, emb_flip = eng.embed(eng.to_device(np.ascontiguousarray(chips_bgr)), n, bool(flip))
        eng.sync()
        e = emb[:n].cpu().numpy()
        return (e, emb_flip[:n].cpu().numpy()) if flip else e
