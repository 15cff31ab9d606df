"""ctypes binding of libpcb200.so (include/pcb200.h).

Follows the reference's own native-binding precedent (person_capture/hdr_preview.py:19-102:
ctypes.CDLL + opaque context pointer).  There is no CPU fallback: if the shared library is
missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCB_LIB") or os.path.join(_HERE, "libpcb200.so")   # PCB_LIB: A/B another build of the library
VAL_LIB_PATH = os.path.join(_HERE, "libpcb200_val.so")   # product + the CUDA-core validation convolution (tests / tools only)

# P-layout padding (csrc/pcb_common.cuh kPadLo / kPad): activations are [n][h + P_PAD][w + P_PAD][cp], image pixel (y, x) at
# [y + P_PAD_LO][x + P_PAD_LO], zeros elsewhere.  load() overwrites these with what the loaded library reports (pcb_layout_pad).
P_PAD_LO = 1
P_PAD = 2

FEAT_DIM = 512
CHIP = 112

OP_CONV, OP_AFFINE, OP_MAXPOOL3S2, OP_AVGPOOL2, OP_UPSAMPLE_ADD, OP_ADD, OP_AFFINE_FLATTEN, OP_FC = range(1, 9)
ACT_NONE, ACT_RELU, ACT_PRELU = 0, 1, 2
MODEL_SCRFD, MODEL_ARCFACE = 0, 1
FIX_NONE, FIX_SCALE, FIX_PADPROBE, FIX_UNPAD = 0, 1, 2, 3
OPF_OUT_F32 = 1

EXPORTS = (
    "pcb_create", "pcb_destroy", "pcb_last_error", "pcb_sync", "pcb_layout_pad", "pcb_set_conv_impl", "pcb_launch_count",
    "pcb_reset_launch_count", "pcb_set_profile", "pcb_profile_read", "pcb_model_load", "pcb_model_get_tensor", "pcb_resize_area", "pcb_resize_linear", "pcb_resize_factor",
    "pcb_detect", "pcb_letterbox", "pcb_decode_nms", "pcb_align", "pcb_embed", "pcb_set_bank", "pcb_match", "pcb_replay",
    "pcb_bank_create", "pcb_bank_destroy", "pcb_bank_offer", "pcb_bank_rows", "pcb_bank_version", "pcb_bank_data",
    "pcb_live_begin", "pcb_live_refresh",
)


class PcbOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("in0", C.c_int32), ("in1", C.c_int32), ("out", C.c_int32),
                ("cin", C.c_int32), ("cout", C.c_int32), ("k", C.c_int32), ("stride", C.c_int32), ("act", C.c_int32),
                ("out2", C.c_int32), ("flags", C.c_int32),
                ("w_off", C.c_int64), ("scale_off", C.c_int64), ("bias_off", C.c_int64), ("slope_off", C.c_int64),
                ("scale2_off", C.c_int64), ("bias2_off", C.c_int64)]


class DetectArgs(C.Structure):
    _fields_ = [("frames_dev", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("S", C.c_int32),
                ("det_thresh", C.c_float), ("rot_deg", C.c_int32), ("pad_replicate", C.c_int32), ("fix_mode", C.c_int32),
                ("fix_scale_inv", C.c_float), ("orig_h", C.c_int32), ("orig_w", C.c_int32), ("min_box_px", C.c_int32),
                ("max_det", C.c_int32),
                ("det_dev", C.c_void_p), ("kps_dev", C.c_void_p), ("raw_count_dev", C.c_void_p),
                ("acc_box_dev", C.c_void_p), ("acc_kps_dev", C.c_void_p), ("acc_score_dev", C.c_void_p),
                ("acc_count_dev", C.c_void_p), ("acc_unfiltered_dev", C.c_void_p)]


class AlignArgs(C.Structure):
    _fields_ = [("frames_dev", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("max_det", C.c_int32),
                ("acc_box_dev", C.c_void_p), ("acc_kps_dev", C.c_void_p), ("acc_score_dev", C.c_void_p),
                ("acc_count_dev", C.c_void_p), ("max_faces", C.c_int32),
                ("face_count_dev", C.c_void_p), ("face_total_dev", C.c_void_p), ("face_frame_dev", C.c_void_p),
                ("face_box_dev", C.c_void_p), ("face_kind_dev", C.c_void_p), ("chips_dev", C.c_void_p),
                ("quality_dev", C.c_void_p)]


REPLAY_META = 10


class ReplayCfg(C.Structure):
    _fields_ = [("enter", C.c_double), ("exit_thr", C.c_double), ("fd_add", C.c_double), ("quality_min", C.c_double),
                ("total_frames", C.c_int64), ("pad", C.c_int64), ("min_len", C.c_int64), ("exit_cool", C.c_int64),
                ("stride", C.c_int32), ("cooldown", C.c_int32), ("fd9_skip", C.c_int32), ("fd9_grace", C.c_int32),
                ("fd9_period", C.c_int32)]


class ReplayState(C.Structure):
    _fields_ = [("frame_idx", C.c_int64), ("last_face_idx", C.c_int64), ("no_face_streak", C.c_int32),
                ("rot_cycle", C.c_int32), ("prescan_rr", C.c_int32), ("trk_active", C.c_int32)]


REPLAY_REFRESH_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int)
REPLAY_FLIP_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int)
REPLAY_ABORTED = 5
BANK_SKIP, BANK_ADDED, BANK_DUP, BANK_REPLACED = 0, 1, 2, 3
BANK_ACTIONS = ("skip", "added", "dup", "replaced")


class BankCfg(C.Structure):
    _fields_ = [("cap", C.c_int32), ("dedup", C.c_double), ("margin", C.c_double), ("wa", C.c_double), ("wd", C.c_double),
                ("wq", C.c_double)]


class ReplayIO(C.Structure):
    _fields_ = [("meta", C.c_void_p), ("frame_idx", C.c_void_p), ("n_samples", C.c_int32), ("n_rows", C.c_int32),
                ("quality", C.c_void_p), ("area", C.c_void_p), ("flip_ready", C.c_void_p),
                ("feat_plain", C.c_void_p), ("feat_flip", C.c_void_p), ("fd_plain", C.c_void_p), ("fd_flip", C.c_void_p),
                ("refresh", REPLAY_REFRESH_CB), ("need_flip", REPLAY_FLIP_CB), ("user", C.c_void_p),
                ("best_out", C.c_void_p), ("skip_out", C.c_void_p), ("active_out", C.c_void_p), ("nfaces_out", C.c_void_p),
                ("spans_out", C.c_void_p), ("max_spans", C.c_int32), ("n_spans_out", C.POINTER(C.c_int32)),
                ("refreshes_out", C.POINTER(C.c_int64)), ("feat_plain_dev", C.c_void_p), ("feat_flip_dev", C.c_void_p)]


class PcbError(RuntimeError):
    pass


_libs = {}


def load(path: str = None):
    """Load libpcb200.so (or the library at `path`); raises if it has not been built (python __graft_entry__.py build)."""
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.isfile(path):
        raise PcbError(f"{path} not found: build it with `make -C person_capture_b200/csrc` "
                       "(there is no CPU fallback for the identity path)")
    lib = C.CDLL(path)
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    lib.pcb_create.restype = vp
    lib.pcb_create.argtypes = [i32, vp]
    lib.pcb_destroy.restype = None
    lib.pcb_destroy.argtypes = [vp]
    lib.pcb_last_error.restype = C.c_char_p
    lib.pcb_last_error.argtypes = [vp]
    lib.pcb_sync.argtypes = [vp]
    lib.pcb_set_conv_impl.argtypes = [vp, i32]
    lib.pcb_launch_count.restype = C.c_longlong
    lib.pcb_launch_count.argtypes = [vp]
    lib.pcb_reset_launch_count.restype = None
    lib.pcb_reset_launch_count.argtypes = [vp]
    lib.pcb_set_profile.argtypes = [vp, i32]
    lib.pcb_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong), i32]
    lib.pcb_model_load.argtypes = [vp, i32, C.POINTER(PcbOp), i32, i32, vp, C.c_size_t, C.POINTER(C.c_int32), i32,
                                   C.POINTER(C.c_float)]
    lib.pcb_model_get_tensor.argtypes = [vp, i32, i32, vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.pcb_resize_area.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32]
    lib.pcb_resize_linear.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32]
    lib.pcb_resize_factor.argtypes = [vp, vp, i32, i32, i32, vp, C.c_double, C.c_double, i32]
    lib.pcb_detect.argtypes = [vp, C.POINTER(DetectArgs)]
    lib.pcb_letterbox.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.pcb_decode_nms.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_float), C.POINTER(DetectArgs), f32]
    lib.pcb_align.argtypes = [vp, C.POINTER(AlignArgs)]
    lib.pcb_embed.argtypes = [vp, vp, i32, vp, vp]
    lib.pcb_set_bank.argtypes = [vp, vp, i32]
    lib.pcb_match.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp]
    lib.pcb_replay.argtypes = [vp, C.POINTER(ReplayCfg), vp, C.POINTER(ReplayIO), C.POINTER(ReplayState)]
    lib.pcb_bank_create.restype = vp
    lib.pcb_bank_create.argtypes = [C.POINTER(BankCfg), vp, i32]
    lib.pcb_bank_destroy.restype = None
    lib.pcb_bank_destroy.argtypes = [vp]
    lib.pcb_bank_offer.argtypes = [vp, vp, C.c_double, C.POINTER(C.c_int32)]
    lib.pcb_bank_rows.argtypes = [vp]
    lib.pcb_bank_version.restype = C.c_longlong
    lib.pcb_bank_version.argtypes = [vp]
    lib.pcb_bank_data.restype = C.POINTER(C.c_float)
    lib.pcb_bank_data.argtypes = [vp]
    lib.pcb_live_begin.argtypes = [vp, vp, i32]
    lib.pcb_live_refresh.argtypes = [vp, vp, i32, i32, i32, i32, C.POINTER(C.POINTER(C.c_float))]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int:
            pass
    global P_PAD_LO, P_PAD
    lo, pad = C.c_int(), C.c_int()
    lib.pcb_layout_pad.restype = None
    lib.pcb_layout_pad(C.byref(lo), C.byref(pad))
    P_PAD_LO, P_PAD = lo.value, pad.value     # follow the library that was actually loaded (PCB_LIB A/B builds)
    _libs[path] = lib
    return lib
