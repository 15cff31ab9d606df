"""Runs the UNMODIFIED reference (`/root/reference/person_capture`) in this container so that golden vectors can be taken
from the reference itself rather than from our restatement of it.

What makes that possible offline: `person_capture/face_embedder.py` imports with the packages this image has (cv2, numpy,
PIL); `person_capture/gui_app.py` only needs `PySide6` to exist as a module (widgets are defined at import time, nothing is
instantiated), so an inert stand-in is registered for it.  The two model sessions the reference would build through ONNX
Runtime / InsightFace (absent here) are handed in by the caller: any object with `.detect(img, input_size=(S, S))` /
`.det_thresh` for SCRFD and `.run(None, {name: blob})` for ArcFace.  Everything between those two calls -- size / rotation
policy, accumulate + min-size, cross-pass NMS, canonicalisation, alignment, eye-roll fallback, quality, flip + sum + normalise,
output assembly, `_fd_min`, `_stream_ref_bank_update`, the `_prescan` state machine, bridge / refine, the cache file -- is the
reference's own code, executed from where it lies (nothing is copied into this repository).

Test infrastructure (tests/golden/make_reference_golden.py is the only caller); `/root/reference` does not exist on the GPU
box, so nothing in tests/ imports this module at run time.
"""
from __future__ import annotations

import importlib
import logging
import sys
import types

REFERENCE_ROOT = "/root/reference"


class _Inert:
    """Instance stand-in: every attribute is another inert object, calling it returns one."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert()

    def connect(self, *a, **k):
        pass

    def emit(self, *a, **k):
        pass


class _InertMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _inert_class(name)


def _inert_class(name):
    def _getattr(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Inert()

    return _InertMeta(name, (object,), {"__init__": lambda self, *a, **k: None, "__getattr__": _getattr})


class _InertModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name == "Slot":
            return lambda *a, **k: (lambda fn: fn)
        if name in ("Signal", "Property"):
            return lambda *a, **k: _Inert()
        cls = _inert_class(name)
        setattr(self, name, cls)
        return cls


def import_reference():
    """-> (face_embedder module, gui_app module) of the reference, imported from REFERENCE_ROOT."""
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "PySide6" not in sys.modules:
        top = _InertModule("PySide6")
        sys.modules["PySide6"] = top
        for sub in ("QtCore", "QtGui", "QtWidgets", "QtMultimedia", "QtMultimediaWidgets", "QtOpenGLWidgets", "QtOpenGL"):
            m = _InertModule("PySide6." + sub)
            sys.modules["PySide6." + sub] = m
            setattr(top, sub, m)
    if "onnxruntime" not in sys.modules:
        # `_get_scrfd_trt` (face_embedder.py:966-987) does `import onnxruntime` before it returns the already-built detector
        # singleton; the module only has to exist (the singleton is the caller's SCRFD stand-in, no session is created)
        sys.modules["onnxruntime"] = _InertModule("onnxruntime")
    logging.disable(logging.CRITICAL)          # the reference logs its (failed) HDR / TRT probing at import
    try:
        fe = importlib.import_module("person_capture.face_embedder")
        ga = importlib.import_module("person_capture.gui_app")
    finally:
        logging.disable(logging.NOTSET)
    return fe, ga


class _ArcIO:
    def __init__(self, name):
        self.name = name


class ArcSessionShim:
    """The slice of ort.InferenceSession `_arcface_encode` touches (face_embedder.py:1369), around fn(blob[n,3,112,112]) -> [n,512]."""

    def __init__(self, fn):
        self.fn = fn
        self.blobs = []

    def get_inputs(self):
        return [_ArcIO("input.1")]

    def get_outputs(self):
        return [_ArcIO("embedding")]

    def get_providers(self):
        return ["CPUExecutionProvider"]

    def run(self, names, feeds):
        x = next(iter(feeds.values()))
        self.blobs.append(x.copy())
        return [self.fn(x)]


def make_reference_embedder(fe_mod, scrfd, arc_fn, conf=0.5):
    """A reference FaceEmbedder whose constructor side effects (downloads, ORT sessions) are skipped: the attributes
    `__init__` would set (face_embedder.py:389-497, 716-725) are set here to the same defaults, and the two sessions are
    the caller's."""
    import torch

    F = fe_mod.FaceEmbedder
    o = object.__new__(F)
    o.detector_backend = "scrfd"
    o.det = None
    o.scrfd = scrfd
    o._scrfd_trt_singleton = scrfd            # what _get_scrfd_trt hands to every detection pass
    o._SCRFD_cls = type(scrfd)                # looked up (never instantiated) before the singleton is returned
    o._scrfd_model_path = None
    o._scrfd_ctx_id = 0
    o._scrfd_fixed_shape = (640, 640)
    o._scrfd_is_trt = False
    o._scrfd_is_cuda = False
    o.insight_app = None
    o.conf = float(conf)
    o.progress = None
    o._torch = torch
    o.trt_lib_dir = None
    o.scrfd_tta_scales = (0.75, 0.60)
    o.scrfd_probe_conf_cap = 0.20
    o.scrfd_edge_pad_frac = 0.06
    o._fast_prescan = False
    o._prescan_rr = 0
    o._prescan_rr_mode = "rr"
    o._prescan_escalate = False
    o._probe_conf = 0.03
    o._high_90 = 1536
    o._high_180 = 1280
    o._prescan_period = 3
    o._prescan_probe_imgsz = 384
    o._prescan_no_upscale_det = True
    o._heavy_cap = 2048
    o._frame_idx = 0
    o._no_face_streak = 0
    o._last_face_idx = -10 ** 9
    o._rot_cycle = 0
    o.rot_adaptive = True
    o.rot_every_n = 12
    o.rot_after_hit_frames = 8
    o.fast_no_face_imgsz = 512
    o.use_arcface = True
    o.device = "cpu"
    o.backend = "arcface"
    o.arc_sess = ArcSessionShim(arc_fn)
    o._arc_fixed_batch = False
    o.arc_input = "input.1"
    o.arc_output = "embedding"
    o._arc_scratch = None
    o._arc_fixed_batch_len = None
    o._arc_feat_dim = 512
    return o
