"""Synthetic weights / inputs for benchmarks and smoke runs (SURVEY.md §8d distributions; numpy PCG64).
There is no network for datasets or checkpoints; shapes follow Potsdam / LoveDA tiles."""
from __future__ import annotations

import math
from typing import Dict

import numpy as np


def msda_state(seed=1234, embed_dim=256, num_heads=8, num_levels=3, num_points=6, offset_std=0.05, attn_std=0.1
               ) -> Dict[str, np.ndarray]:
    """Non-trivial MSDeformableAttention weights with Paddle state-dict keys / layouts ([in, out]).  The reference
    init zeroes sampling_offsets.weight and attention_weights.* (t_e_d.py:47-58), which would make the sampled
    positions input-independent; benchmarks must exercise the data-dependent path."""
    rng = np.random.Generator(np.random.PCG64(seed))
    C, tp = embed_dim, num_heads * num_levels * num_points
    xav = math.sqrt(6.0 / (C + C))
    thetas = np.arange(num_heads, dtype=np.float32) * np.float32(2.0 * math.pi / num_heads)
    grid = np.stack([np.cos(thetas), np.sin(thetas)], -1)
    grid = grid / np.abs(grid).max(-1, keepdims=True)
    grid = np.tile(grid.reshape(num_heads, 1, 1, 2), (1, num_levels, num_points, 1))
    grid = grid * np.arange(1, num_points + 1, dtype=np.float32).reshape(1, 1, -1, 1)
    f = np.float32
    return {
        "sampling_offsets.weight": (rng.standard_normal((C, tp * 2)) * offset_std).astype(f),
        "sampling_offsets.bias": grid.reshape(-1).astype(f),
        "attention_weights.weight": (rng.standard_normal((C, tp)) * attn_std).astype(f),
        "attention_weights.bias": rng.uniform(-0.1, 0.1, (tp,)).astype(f),
        "value_proj.weight": rng.uniform(-xav, xav, (C, C)).astype(f),
        "value_proj.bias": rng.uniform(-0.1, 0.1, (C,)).astype(f),
        "output_proj.weight": rng.uniform(-xav, xav, (C, C)).astype(f),
        "output_proj.bias": rng.uniform(-0.1, 0.1, (C,)).astype(f),
    }


def encoder_reference_points(shapes) -> np.ndarray:
    """[1, Lv, L, 2] pixel-centre reference points (t_e_d.py:213-228 with valid_ratios == 1)."""
    pts = []
    for (H, W) in shapes:
        ys = (np.arange(H, dtype=np.float32) + 0.5) / np.float32(H)
        xs = (np.arange(W, dtype=np.float32) + 0.5) / np.float32(W)
        yy, xx = np.meshgrid(ys, xs, indexing="ij")
        pts.append(np.stack([xx.reshape(-1), yy.reshape(-1)], -1))
    ref = np.concatenate(pts, 0)[None, :, None, :]
    return np.ascontiguousarray(np.broadcast_to(ref, (1, ref.shape[1], len(shapes), 2))).astype(np.float32)


def level_shapes(tile: int):
    """C3..C5 feature-map shapes of a tile x tile input (strides 8/16/32; paddle_EMRT.py:254)."""
    return [(tile // 8, tile // 8), (tile // 16, tile // 16), (tile // 32, tile // 32)]


def encoder_layer_state(seed=1234, C=256, ffn=1024, heads=8, levels=3, points=6) -> Dict[str, np.ndarray]:
    """Non-trivial TransformerEncoderLayer weights with the reference's state-dict keys and Paddle layouts
    (transformer_encoder_decoder.py:109-147): Linear [in,out], Conv2D [out,in,3,3], norm affine gamma ~ U(0.5,1.5),
    beta ~ N(0,0.1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f = np.float32
    p: Dict[str, np.ndarray] = {}
    for k, v in msda_state(int(rng.integers(1 << 30)), C, heads, levels, points).items():
        p["self_attn." + k] = v

    def norm(name):
        p[name + ".weight"] = rng.uniform(0.5, 1.5, size=(C,)).astype(f)
        p[name + ".bias"] = (rng.standard_normal((C,)) * 0.1).astype(f)

    def linear(name, i, o):
        b = math.sqrt(6.0 / (i + o))
        p[name + ".weight"] = rng.uniform(-b, b, (i, o)).astype(f)
        p[name + ".bias"] = rng.uniform(-0.1, 0.1, (o,)).astype(f)
    norm("norm1")
    linear("linear1", C, ffn)
    linear("linear2", ffn, C)
    norm("norm2")
    for l in range(3):
        b = math.sqrt(6.0 / (2 * 9 * C))
        p[f"conv{l}.0.weight"] = rng.uniform(-b, b, (C, C, 3, 3)).astype(f)
        norm(f"conv{l}.1")
    return p


def encoder_decoder_state(seed=1234, C=256, ffn=1024, heads=8, levels=3, points=6, num_enc=4, num_dec=2,
                          in_channels=(512, 1024, 2048), num_queries=110) -> Dict[str, np.ndarray]:
    """Non-trivial weights for the whole EncoderDecoder (transformer_encoder_decoder.py:337-407) with the reference's
    state-dict keys: encoder.layers.{i}.*, decoder.layers.{i}.{self_attn.in_proj_weight [C,3C], self_attn.in_proj_bias,
    self_attn.out_proj.*, norm1-3, cross_attn.*, linear1-2}, level_embed / tgt_embed / query_pos_embed .weight,
    reference_points.*, input_proj.{l}.{0,1}.*.  Same distributions as the per-module generators above."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f = np.float32
    p: Dict[str, np.ndarray] = {}

    def norm(name, c=C):
        p[name + ".weight"] = rng.uniform(0.5, 1.5, size=(c,)).astype(f)
        p[name + ".bias"] = (rng.standard_normal((c,)) * 0.1).astype(f)

    def linear(name, i, o):
        b = math.sqrt(6.0 / (i + o))
        p[name + ".weight"] = rng.uniform(-b, b, (i, o)).astype(f)
        p[name + ".bias"] = rng.uniform(-0.1, 0.1, (o,)).astype(f)

    for i in range(num_enc):
        for k, v in encoder_layer_state(int(rng.integers(1 << 30)), C, ffn, heads, levels, points).items():
            p[f"encoder.layers.{i}.{k}"] = v
    for i in range(num_dec):
        pre = f"decoder.layers.{i}"
        b = math.sqrt(6.0 / (C + 3 * C))
        p[pre + ".self_attn.in_proj_weight"] = rng.uniform(-b, b, (C, 3 * C)).astype(f)
        p[pre + ".self_attn.in_proj_bias"] = rng.uniform(-0.1, 0.1, (3 * C,)).astype(f)
        linear(pre + ".self_attn.out_proj", C, C)
        norm(pre + ".norm1")
        for k, v in msda_state(int(rng.integers(1 << 30)), C, heads, levels, points).items():
            p[f"{pre}.cross_attn.{k}"] = v
        norm(pre + ".norm2")
        linear(pre + ".linear1", C, ffn)
        linear(pre + ".linear2", ffn, C)
        norm(pre + ".norm3")
    p["level_embed.weight"] = rng.standard_normal((levels, C)).astype(f)
    p["tgt_embed.weight"] = rng.standard_normal((num_queries, C)).astype(f)
    p["query_pos_embed.weight"] = rng.standard_normal((num_queries, C)).astype(f)
    linear("reference_points", C, 2)
    for l, cin in enumerate(in_channels):
        b = math.sqrt(6.0 / (cin + C))
        p[f"input_proj.{l}.0.weight"] = rng.uniform(-b, b, (C, cin, 1, 1)).astype(f)
        p[f"input_proj.{l}.0.bias"] = rng.uniform(-0.1, 0.1, (C,)).astype(f)
        norm(f"input_proj.{l}.1")
    return p
