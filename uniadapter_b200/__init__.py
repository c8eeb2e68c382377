"""uniadapter_b200 — sm_100a implementation of Uni-Adapter's per-sample test-time hot path.

Point tokenizer (FPS, kNN / ball-query grouping, centre normalisation), zero-shot cosine head, DOTA / MODE-DOTA cache
step and their fusion, behind the reference's Python call signatures. All compute goes through libua_b200.so
(include/ua_b200.h); there is no CPU or PyTorch fallback for those ops.
"""
from . import _lib  # noqa: F401
from .tokenizer import (Group, ball_group, farthest_point_sample, fps, fps_sample, fps_uni3d,  # noqa: F401
                        furthest_point_sample, gather_operation, knn_group, knn_point, query_ball_point,
                        sample_and_group)
from .head import get_logits_wrapper, softmax_entropy, zero_shot_head  # noqa: F401
from .dota import DOTA  # noqa: F401
from .dota_mixture import DOTA_mix  # noqa: F401
from .fusion import fuse_logits  # noqa: F401

__version__ = "0.1.0"
