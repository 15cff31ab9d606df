"""The slice of the reference SessionConfig that parameterises the identity path.

Field names and defaults follow person_capture/gui_app.py SessionConfig (:295-638): the
prescan_* block (:553-594), face_* (:467-477), rotation strategy (:523-527).  Everything else
in SessionConfig (crop composition, HDR export, UI) is out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Tuple


@dataclass
class PrescanParams:
    # inputs that enter the cache key (gui_app.py:787-841)
    video: str = ""
    ref: str = ""
    # identity thresholds
    face_thresh: float = 0.45
    face_quality_min: float = 70.0
    face_det_conf: float = 0.5
    face_model: str = "scrfd_10g_bnkps"
    face_fullframe_imgsz: int = 1408
    clip_face_backbone: str = "ViT-L-14"
    clip_face_pretrained: str = "laion2b_s32b_b82k"
    use_arcface: bool = True
    # main-pass identity sites (gui_app.py:308, 419-421, 468, 471, 477, 522)
    frame_stride: int = 2
    lock_face_roi_enable: bool = True
    lock_face_roi_pad: float = 1.25
    lock_face_roi_max_misses: int = 8
    face_fullframe_cadence: int = 12
    face_visible_uses_quality: bool = True
    learn_bank_runtime: bool = False
    face_fullframe_when_missed: bool = True
    # rotation strategy of the main pass
    rot_adaptive: bool = True
    rot_every_n: int = 12
    rot_after_hit_frames: int = 8
    fast_no_face_imgsz: int = 512
    # pre-scan
    prescan_stride: int = 24
    prescan_max_width: int = 416
    prescan_decode_max_w: int = 384
    prescan_face_conf: float = 0.5
    prescan_fd_enter: float = 0.45
    prescan_fd_add: float = 0.22
    prescan_fd_exit: float = 0.52
    prescan_add_cooldown_samples: int = 5
    prescan_rot_probe_period: int = 3
    prescan_probe_imgsz: int = 512
    prescan_no_upscale_det: bool = True
    prescan_probe_conf: float = 0.03
    prescan_heavy_90: int = 1536
    prescan_heavy_180: int = 1280
    prescan_min_segment_sec: float = 1.0
    prescan_pad_sec: float = 1.5
    prescan_bridge_gap_sec: float = 1.0
    prescan_exit_cooldown_sec: float = 0.5
    prescan_boundary_refine_sec: float = 0.75
    prescan_refine_stride_min: int = 3
    prescan_trim_pad: bool = True
    prescan_skip_trailing_refine: bool = True
    prescan_refine_budget_sec: float = 1.5
    prescan_bank_max: int = 64
    prescan_diversity_dedup_cos: float = 0.968
    prescan_replace_margin: float = 0.010
    prescan_fd9_skip: bool = True
    prescan_fd9_grace: int = 1
    prescan_fd9_probe_period: int = 2
    prescan_weights: Tuple[float, float, float] = (0.70, 0.25, 0.05)
    prescan_cache_mode: str = "auto"
    prescan_cache_dir: str = "prescan_cache"

    def to_dict(self):
        return asdict(self)
