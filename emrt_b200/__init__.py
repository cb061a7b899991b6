"""emrt_b200 — B200-native (sm_100a) implementation of EMRT's data-parallel hot path.

Public surface (mirrors the reference's Python surface for this path):
  MSDeformableAttention, deformable_attention_core_func   (src/models/EMRT_utils/*)
  slide_inference, ss_inference                           (src/api/infer.py)
  calculate_area                                          (src/utils/metrics.py)
Everything computes in libemrt_b200.so through the C ABI in include/emrt_b200.h; there is no CPU fallback.
"""
from . import _lib
from ._lib import EmrtError
from .msda import MSDeformableAttention, deformable_attention_core_func, shapes_to_host
from .infer import slide_inference, ss_inference, ss_inference_eval, plan_windows, window_origins
from .ops import calculate_area
from .refpoints import get_reference_points
from .encoder import TransformerEncoderLayer, TransformerEncoder
from .decoder import MultiHeadAttention, TransformerDecoderLayer, TransformerDecoder, EncoderDecoder
from .sharding import shard_range, shard_images, shard_scene_rows

__all__ = ["MSDeformableAttention", "deformable_attention_core_func", "slide_inference", "ss_inference", "ss_inference_eval",
           "calculate_area", "plan_windows", "window_origins", "shard_range", "shard_images", "shard_scene_rows",
           "EmrtError", "shapes_to_host", "get_reference_points", "TransformerEncoderLayer", "TransformerEncoder", "MultiHeadAttention",
           "TransformerDecoderLayer", "TransformerDecoder", "EncoderDecoder"]
