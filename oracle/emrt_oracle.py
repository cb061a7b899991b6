"""CPU restatement (torch-CPU / numpy) of the EMRT hot path.  TEST INFRASTRUCTURE ONLY.

PARITY PINNED TO THE REFERENCE'S SOURCES, NOT TO PADDLE'S KERNELS.  The reference ships no tests / golden vectors /
fixtures, and its arithmetic is delegated to the un-vendored third-party package ``paddlepaddle`` (README.md:16
"2.1.0+", README.md:29 ``paddlepaddle-gpu==2.1.2``; not pinned in requirements.txt), which cannot be installed in
this image (no wheel, no network, needs Python <= 3.8).  What pins this restatement:
  (o)  the reference's own, unmodified Python sources executed in this container on a torch-CPU mapping of the
       Paddle operators they call (oracle/paddle_on_torch.py + oracle/run_reference.py): tests/test_reference_pin.py
       checks every function here against vectors generated that way (tests/golden/ref_*.npz, generator
       tests/golden/make_reference_vectors.py) and against live runs of the reference on further shapes;
  (i)  two independent formulations of every sampled op (library ``grid_sample`` / ``interpolate``
       composition that mirrors the reference line by line vs. explicit closed-form corner loops),
  (ii) analytic cases (zero offsets, all-outside samples, linearity, uniform attention),
  (iii) float64 evaluation as the arbiter for fp32 / bf16 tolerances.
Still assumed (from the Paddle API documentation, uncheckable without Paddle): ``nn.Linear`` is ``y = x @ W + b``
with ``W`` stored ``[in, out]``; LayerNorm/GroupNorm eps 1e-5; GELU exact (erf); ``F.grid_sample`` grid last
dim is (x, y); ``F.interpolate(align_corners=False)`` uses half-pixel centres (align_mode=0);
``paddle.argmax`` returns the first maximal index.

All file:line citations are relative to /root/reference/semantic_segmentation/.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

__all__ = [
    "level_tables", "gather_corner_loop", "deformable_attention_core_func", "msda_reset_parameters",
    "make_msda_params", "msda_forward", "msda_intermediates", "encoder_reference_points",
    "upsample2x", "upsample2x_loop", "interpolate_bilinear", "window_origins", "slide_inference",
    "ss_inference_tail", "ss_inference", "calculate_area", "position_embedding_sine",
    "multi_head_attention", "encoder_layer_forward", "decoder_layer_forward", "encoder_decoder_forward",
    "make_encoder_decoder_params", "rng_normal", "rng_uniform", "kernel_storage_rounding",
]


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def level_tables(shapes: Sequence[Tuple[int, int]]):
    """(H_l, W_l) list -> (level_start, Lv).  Mirrors the split at src/models/EMRT_utils/utils.py:77."""
    start, acc = [], 0
    for h, w in shapes:
        start.append(acc)
        acc += int(h) * int(w)
    return start, acc


# ----------------------------------------------------------------------------------------------
# storage-rounding mode: the restatement evaluated with the B200 path's STORAGE formats
# ----------------------------------------------------------------------------------------------
# The bf16 path keeps every tensor that crosses HBM in a 16-bit format: activations bf16, sampling offsets
# (in pixels of their level) and softmax weights fp16; all arithmetic between two stores is fp32.  Inside
# `kernel_storage_rounding()` the functions below apply exactly those roundings — and nothing else — at the points where
# the kernels store, while still computing in the caller's dtype (float64 in the tests).  Two uses:
#   * kernel vs this = the kernels' OWN error (accumulation order, exp / erf approximations), asserted tight (tests/);
#   * this vs the exact evaluation = what the storage formats cost at depth, measured on the CPU alone
#     (tests/test_oracle.py) — the part of the bf16 tolerance that no kernel can remove.
_ROUND = {"on": False}


class kernel_storage_rounding:
    def __enter__(self):
        self.prev = _ROUND["on"]
        _ROUND["on"] = True
        return self

    def __exit__(self, *exc):
        _ROUND["on"] = self.prev
        return False


def _store(x, kind="act"):
    """Identity, or (storage-rounding mode) x rounded to the format the kernels store this tensor in."""
    if not _ROUND["on"]:
        return x
    fmt = torch.bfloat16 if kind == "act" else torch.float16
    return x.to(fmt).to(x.dtype)


def rng_normal(rng: np.random.Generator, shape, std=1.0, dtype=np.float32):
    return (rng.standard_normal(size=shape) * std).astype(dtype)


def rng_uniform(rng: np.random.Generator, shape, bound, dtype=np.float32):
    return rng.uniform(-bound, bound, size=shape).astype(dtype)


# ----------------------------------------------------------------------------------------------
# a3: the multiscale bilinear sampling-gather
# ----------------------------------------------------------------------------------------------
def deformable_attention_core_func(value, value_spatial_shapes, sampling_locations, attention_weights):
    """Line-by-line restatement of src/models/EMRT_utils/utils.py:64-97 with torch ops.

    value [bs, Lv, M, D]; value_spatial_shapes [(H,W)]*L; sampling_locations [bs, Lq, M, L, P, 2] (x,y in [0,1]);
    attention_weights [bs, Lq, M, L, P]  ->  [bs, Lq, M*D]
    """
    value = torch.as_tensor(value)
    sampling_locations = torch.as_tensor(sampling_locations)
    attention_weights = torch.as_tensor(attention_weights)
    if _ROUND["on"] and sampling_locations.shape[1] == value.shape[1]:
        # encoder self-attention geometry (one query per value pixel): the window-staged kernels' corner-weight rounding;
        # the decoder's kernel (msda_gather_v1.cu) keeps fp32 corner weights — nothing to round there
        return _gather_bf16_corner_weights(value, value_spatial_shapes, sampling_locations, attention_weights)
    bs, Len_v, n_head, c = value.shape
    _, Len_q, _, n_levels, n_points, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in value_spatial_shapes]
    value_list = value.split([h * w for h, w in shapes], dim=1)            # utils.py:77
    sampling_grids = 2 * sampling_locations - 1                             # utils.py:79
    sampling_value_list = []
    for level, (h, w) in enumerate(shapes):                                 # utils.py:82
        value_l_ = value_list[level].flatten(2).permute(0, 2, 1).reshape(bs * n_head, c, h, w)   # :83
        sampling_grid_l_ = sampling_grids[:, :, :, level].permute(0, 2, 1, 3, 4).flatten(0, 1)   # :84
        sampling_value_l_ = F.grid_sample(value_l_, sampling_grid_l_, mode="bilinear",
                                          padding_mode="zeros", align_corners=False)              # :87-88
        sampling_value_list.append(sampling_value_l_)
    attention_weights = attention_weights.permute(0, 2, 1, 3, 4).reshape(
        bs * n_head, 1, Len_q, n_levels * n_points)                                               # :91-92
    output = (torch.stack(sampling_value_list, dim=-2).flatten(-2) * attention_weights).sum(-1)  # :94
    output = output.reshape(bs, n_head * c, Len_q)                                                # :95
    return output.permute(0, 2, 1).contiguous()                                                   # :97


def _gather_bf16_corner_weights(value, shapes, loc, attn):
    """Storage-rounding mode of the gather: the closed form of `gather_corner_loop` with each of the four corner weights
    ((1-fx)*aw)*(1-fy) ... rounded to bf16 — the B200 gather kernels multiply bf16 values by bf16 corner weights into an
    fp32 accumulator (`fma.rn.f32.bf16`, msda_win_common.cuh) — everything else in the caller's dtype."""
    B, Lv, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    shapes = [(int(h), int(w)) for h, w in shapes]
    start, total = level_tables(shapes)
    out = torch.zeros((B, Lq, M, D), dtype=value.dtype)
    bi = torch.arange(B)[:, None, None]
    mi = torch.arange(M)[None, None, :]
    for l, (H, W) in enumerate(shapes):
        for p in range(P):
            x = loc[:, :, :, l, p, 0] * W - 0.5
            y = loc[:, :, :, l, p, 1] * H - 0.5
            x0, y0 = torch.floor(x), torch.floor(y)
            fx, fy = x - x0, y - y0
            aw = attn[:, :, :, l, p]
            for dy, dx in ((0, 0), (0, 1), (1, 0), (1, 1)):
                xi, yi = (x0 + dx).long(), (y0 + dy).long()
                w = _store(((fx if dx else 1.0 - fx) * aw) * (fy if dy else 1.0 - fy))
                inside = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
                idx = start[l] + yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)
                out += torch.where(inside, w, torch.zeros_like(w))[..., None] * value[bi, idx, mi]
    return out.reshape(B, Lq, M * D)


def gather_corner_loop(value, shapes, loc, attn):
    """Independent closed-form restatement of utils.py:64-97, numpy float64.

    out[b,q,m*D+d] = sum_l sum_p aw * sum_{corner} w_corner * value[b, start_l + y*W_l + x, m, d]
    with x = loc_x*W_l - 0.5, y = loc_y*H_l - 0.5 and each corner contributing only if inside the map
    (grid_sample bilinear / zeros / align_corners=False).
    """
    value = np.asarray(value, dtype=np.float64)
    loc = np.asarray(loc, dtype=np.float64)
    attn = np.asarray(attn, dtype=np.float64)
    B, Lv, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    start, total = level_tables(shapes)
    assert total == Lv
    out = np.zeros((B, Lq, M, D), dtype=np.float64)
    bi = np.arange(B)[:, None, None]
    mi = np.arange(M)[None, None, :]
    for l, (H, W) in enumerate(shapes):
        for p in range(P):
            x = loc[:, :, :, l, p, 0] * W - 0.5
            y = loc[:, :, :, l, p, 1] * H - 0.5
            x0 = np.floor(x)
            y0 = np.floor(y)
            lx, ly = x - x0, y - y0
            aw = attn[:, :, :, l, p]
            for dy, dx in ((0, 0), (0, 1), (1, 0), (1, 1)):
                xi = (x0 + dx).astype(np.int64)
                yi = (y0 + dy).astype(np.int64)
                w = (lx if dx else 1.0 - lx) * (ly if dy else 1.0 - ly) * aw
                inside = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
                idx = start[l] + np.clip(yi, 0, H - 1) * W + np.clip(xi, 0, W - 1)
                v = value[bi, idx, mi]                     # [B,Lq,M,D]
                out += np.where(inside, w, 0.0)[..., None] * v
    return out.reshape(B, Lq, M * D)


# ----------------------------------------------------------------------------------------------
# a1/a2: MSDeformableAttention
# ----------------------------------------------------------------------------------------------
def msda_reset_parameters(embed_dim=256, num_heads=8, num_levels=3, num_points=6):
    """The reference initial sampling_offsets bias (transformer_encoder_decoder.py:47-55).  float32 [M*L*P*2]."""
    thetas = np.arange(num_heads, dtype=np.float32) * np.float32(2.0 * math.pi / num_heads)
    grid = np.stack([np.cos(thetas), np.sin(thetas)], -1)
    grid = grid / np.abs(grid).max(-1, keepdims=True)
    grid = np.tile(grid.reshape(num_heads, 1, 1, 2), (1, num_levels, num_points, 1))
    grid = grid * np.arange(1, num_points + 1, dtype=np.float32).reshape(1, 1, -1, 1)
    return grid.reshape(-1).astype(np.float32)


def make_msda_params(seed=1234, embed_dim=256, num_heads=8, num_levels=3, num_points=6,
                     offset_std=0.05, attn_std=0.1) -> Dict[str, np.ndarray]:
    """Synthetic, data-dependent weights: non-zero sampling_offsets / attention_weights
    weights so the data-dependent path is exercised (the reference init, t_e_d.py:47-58, zeroes them).
    Keys and layouts are the Paddle state-dict ones: Linear weight is [in, out]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    C, tp = embed_dim, num_heads * num_levels * num_points
    xav = lambda i, o: math.sqrt(6.0 / (i + o))
    return {
        "sampling_offsets.weight": rng_normal(rng, (C, tp * 2), offset_std),
        "sampling_offsets.bias": msda_reset_parameters(C, num_heads, num_levels, num_points),
        "attention_weights.weight": rng_normal(rng, (C, tp), attn_std),
        "attention_weights.bias": rng_uniform(rng, (tp,), 0.1),
        "value_proj.weight": rng_uniform(rng, (C, C), xav(C, C)),
        "value_proj.bias": rng_uniform(rng, (C,), 0.1),
        "output_proj.weight": rng_uniform(rng, (C, C), xav(C, C)),
        "output_proj.bias": rng_uniform(rng, (C,), 0.1),
    }


def _t(x, dtype):
    return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(dtype)


def msda_intermediates(params, query, reference_points, value, shapes, value_mask=None,
                       num_heads=8, num_points=6, dtype=torch.float32, query_pos=None):
    """Steps (1)-(4) of MSDeformableAttention.forward (transformer_encoder_decoder.py:79-102).
    Returns (value_proj'd value [B,Lv,M,D], sampling_locations [B,Lq,M,L,P,2], attention [B,Lq,M,L,P])."""
    q = _t(query, dtype)
    v = _t(value, dtype)
    ref = _t(reference_points, dtype)
    P_ = {k: _t(w, dtype) for k, w in params.items()}
    bs, Len_q = q.shape[:2]
    Len_v = v.shape[1]
    L = len(shapes)
    assert sum(int(h) * int(w) for h, w in shapes) == Len_v                           # :81
    C = q.shape[-1]
    v = v @ P_["value_proj.weight"] + P_["value_proj.bias"]                            # :83
    if value_mask is not None:
        v = v * _t(value_mask, dtype).unsqueeze(-1)                                    # :84-86
    v = _store(v).reshape(bs, Len_v, num_heads, C // num_heads)                        # :88
    if query_pos is None:
        off = q @ P_["sampling_offsets.weight"] + P_["sampling_offsets.bias"]
        aw = q @ P_["attention_weights.weight"] + P_["attention_weights.bias"]
    else:       # (query + pos) W + b evaluated as query W + (pos W + b): the second term is a stored fp16 table (msda.py)
        # the table is computed once from the fp32 master weights (keys "<name>#fp32" when the caller rounded the matrices)
        qp = _t(query_pos, dtype)
        w_off = P_.get("sampling_offsets.weight#fp32", P_["sampling_offsets.weight"])
        w_aw = P_.get("attention_weights.weight#fp32", P_["attention_weights.weight"])
        off = q @ P_["sampling_offsets.weight"] + _store(qp @ w_off + P_["sampling_offsets.bias"], "f16")
        aw = q @ P_["attention_weights.weight"] + _store(qp @ w_aw + P_["attention_weights.bias"], "f16")
    off = _store(off, "f16").reshape(bs, Len_q, num_heads, L, num_points, 2)           # :89-90 (stored in pixels, fp16)
    aw = aw.reshape(bs, Len_q, num_heads, L * num_points)                              # :92-93
    aw = _store(F.softmax(aw, -1), "f16").reshape(bs, Len_q, num_heads, L, num_points)  # :95-96
    normalizer = torch.tensor([[float(w), float(h)] for h, w in shapes], dtype=dtype).reshape(
        1, 1, 1, L, 1, 2)                                                              # :98-99  (flip -> (W,H))
    loc = ref.reshape(bs, Len_q, 1, L, 1, 2) + off / normalizer                        # :101-102
    return v, loc, aw


def msda_forward(params, query, reference_points, value, shapes, value_mask=None,
                 num_heads=8, num_points=6, dtype=torch.float32, query_pos=None):
    """MSDeformableAttention.forward (transformer_encoder_decoder.py:65-107).  `query_pos` (only meaningful in
    storage-rounding mode): `query` is then the un-embedded query and with_pos_embed is evaluated the way the kernels do."""
    v, loc, aw = msda_intermediates(params, query, reference_points, value, shapes, value_mask,
                                    num_heads, num_points, dtype, query_pos)
    out = _store(deformable_attention_core_func(v, shapes, loc, aw))                    # :104
    return out @ _t(params["output_proj.weight"], dtype) + _t(params["output_proj.bias"], dtype)   # :106


def encoder_reference_points(shapes, batch=1, dtype=torch.float32):
    """TransformerEncoder.get_reference_points with valid_ratios == 1
    (transformer_encoder_decoder.py:213-228; valid_ratios are ones because EMRT.forward never passes a mask,
    paddle_EMRT.py:265, t_e_d.py:442).  -> [batch, Lv, L, 2] (x, y)."""
    pts = []
    L = len(shapes)
    for (H, W) in shapes:
        ys = torch.linspace(0.5, H - 0.5, H, dtype=dtype)
        xs = torch.linspace(0.5, W - 0.5, W, dtype=dtype)
        ref_y, ref_x = torch.meshgrid(ys, xs, indexing="ij")
        ref_y = ref_y.flatten() / H
        ref_x = ref_x.flatten() / W
        pts.append(torch.stack((ref_x, ref_y), -1))
    ref = torch.cat(pts, 0)[None, :, None, :]            # [1, Lv, 1, 2]
    return ref.expand(batch, -1, L, -1).contiguous()


# ----------------------------------------------------------------------------------------------
# a5: UpHead tail;  a6: slide_inference;  a7: ss_inference tail;  a8: calculate_area
# ----------------------------------------------------------------------------------------------
def interpolate_bilinear(x, size):
    """F.interpolate(x, size, mode='bilinear', align_corners=False) (paddle_EMRT.py:179-180, infer.py:151)."""
    return F.interpolate(torch.as_tensor(x), size=tuple(int(s) for s in size), mode="bilinear", align_corners=False)


def upsample2x(x):
    """UpHead.forward last line (paddle_EMRT.py:178-180): x2 bilinear, align_corners=False."""
    x = torch.as_tensor(x)
    return interpolate_bilinear(x, (2 * x.shape[-2], 2 * x.shape[-1]))


def upsample2x_loop(x):
    """Independent closed form of the x2 half-pixel bilinear: src = (dst + 0.5)/2 - 0.5 clamped at 0,
    taps (floor, min(floor+1, n-1)).  numpy float64."""
    x = np.asarray(x, dtype=np.float64)
    h, w = x.shape[-2:]

    def taps(n_out, n_in):
        s = np.maximum((np.arange(n_out) + 0.5) / 2.0 - 0.5, 0.0)
        i0 = np.floor(s).astype(np.int64)
        i1 = np.minimum(i0 + 1, n_in - 1)
        f = s - i0
        return i0, i1, f
    y0, y1, fy = taps(2 * h, h)
    x0, x1, fx = taps(2 * w, w)
    top = x[..., y0, :][..., :, x0] * (1 - fx) + x[..., y0, :][..., :, x1] * fx
    bot = x[..., y1, :][..., :, x0] * (1 - fx) + x[..., y1, :][..., :, x1] * fx
    return top * (1 - fy)[:, None] + bot * fy[:, None]


def window_origins(size: int, crop: int, stride: int) -> List[int]:
    """Window origins along one axis (src/api/infer.py:43-44, 52-59)."""
    n = max(size - crop + stride - 1, 0) // stride + 1
    out = []
    for r in range(n):
        h1 = r * stride
        if h1 >= size:
            continue
        h2 = min(h1 + crop, size)
        out.append(max(h2 - crop, 0))
    return out


def slide_inference(model: Callable, imgs: Sequence[torch.Tensor], crop_size, stride_size, num_classes):
    """src/api/infer.py:22-80.  ``model(batch)[0]`` -> logits [n, nc, h, w].  crop/stride are (w, h)."""
    batch_size = len(imgs)
    h_img = [img.shape[-2] for img in imgs]
    w_img = [img.shape[-1] for img in imgs]
    max_h, max_w = max(h_img), max(w_img)
    w_crop, h_crop = crop_size
    w_stride, h_stride = stride_size
    rows = max(max_h - h_crop + h_stride - 1, 0) // h_stride + 1
    cols = max(max_w - w_crop + w_stride - 1, 0) // w_stride + 1
    count = torch.zeros([batch_size, 1, max_h, max_w])
    final_logit = torch.zeros([batch_size, num_classes, max_h, max_w])
    for r in range(rows):
        for c in range(cols):
            batch_list, loc_list = [], []
            for i, img in enumerate(imgs):
                h1, w1 = r * h_stride, c * w_stride
                if h1 >= img.shape[-2] or w1 >= img.shape[-1]:
                    continue
                h2 = min(h1 + h_crop, img.shape[-2])
                w2 = min(w1 + w_crop, img.shape[-1])
                h1 = max(h2 - h_crop, 0)
                w1 = max(w2 - w_crop, 0)
                loc_list.append((i, h1, w1, h2, w2))
                batch_list.append(img[:, h1:h2, w1:w2].unsqueeze(0))
            if not batch_list:
                continue
            logits = model(torch.cat(batch_list, 0))[0]
            for i in range(len(batch_list)):
                idx, h1, w1, h2, w2 = loc_list[i]
                final_logit[idx, :, h1:h2, w1:w2] += logits[i]
                count[idx, :, h1:h2, w1:w2] += 1
    out = []
    for i in range(batch_size):
        h, w = imgs[i].shape[-2:]
        out.append(final_logit[i:i + 1, :, :h, :w] / count[i:i + 1, :, :h, :w])
    return out


def ss_inference_tail(logit, shape):
    """src/api/infer.py:150-154: resize -> softmax(axis=1) -> argmax(axis=1, keepdim, int32)."""
    logit = interpolate_bilinear(logit, shape)
    prob = F.softmax(logit, dim=1)
    return torch.argmax(prob, dim=1, keepdim=True).to(torch.int32)


def ss_inference(model, img, ori_shape, is_slide, base_size, stride_size, crop_size, num_classes,
                 rescale_from_ori=False):
    """src/api/infer.py:82-157 (is_slide=True branch and the plain branch)."""
    if not is_slide:
        if not isinstance(img, (list, tuple)):
            raise TypeError("The type of img must be one of collections.abc.Sequence")
        if len(img) != 1:
            raise ValueError("batch_size should be set to 1 while is_slide is False")
        logits = model(img[0])
        if not isinstance(logits, (list, tuple)):
            raise TypeError("The type of logits must be one of collections.abc.Sequence")
        logit_list = [logits[0]]
    else:
        logit_list = slide_inference(model, img, crop_size, stride_size, num_classes)
    if ori_shape is None:
        return logit_list
    return [ss_inference_tail(l, ori_shape[i]) for i, l in enumerate(logit_list)]


def calculate_area(pred, label, num_classes, ignore_index=255):
    """src/utils/metrics.py:20-69 -> (intersect_area, pred_area, label_area), int64 [num_classes]."""
    pred = np.asarray(pred).astype(np.int64).reshape(-1)
    label = np.asarray(label).astype(np.int64).reshape(-1)
    if pred.shape != label.shape:
        raise ValueError("Shape of `pred` and `label should be equal")
    mask = label != ignore_index
    pred = (pred + 1) * mask
    label = (label + 1) * mask
    pa = np.array([(pred == i + 1).sum() for i in range(num_classes)], dtype=np.int64)
    la = np.array([(label == i + 1).sum() for i in range(num_classes)], dtype=np.int64)
    ia = np.array([((pred == i + 1) & (label == i + 1)).sum() for i in range(num_classes)], dtype=np.int64)
    return ia, pa, la


# ----------------------------------------------------------------------------------------------
# "next" rows: the rest of EncoderDecoder
# ----------------------------------------------------------------------------------------------
def position_embedding_sine(h, w, num_pos_feats=128, temperature=10000.0, offset=-0.5, eps=1e-6,
                            scale=2 * math.pi, dtype=torch.float32):
    """PositionEmbedding.forward, sine / normalize=True, all-ones mask (position_encoding.py:51-75).
    -> [h*w, 2*num_pos_feats] in the flattened-token layout used at t_e_d.py:448."""
    y_embed = torch.arange(1, h + 1, dtype=dtype)[:, None].expand(h, w)
    x_embed = torch.arange(1, w + 1, dtype=dtype)[None, :].expand(h, w)
    y_embed = (y_embed + offset) / (y_embed[-1:, :] + eps) * scale
    x_embed = (x_embed + offset) / (x_embed[:, -1:] + eps) * scale
    dim_t = 2 * (torch.arange(num_pos_feats) // 2).to(dtype)
    dim_t = temperature ** (dim_t / num_pos_feats)
    pos_x = x_embed[..., None] / dim_t
    pos_y = y_embed[..., None] / dim_t
    pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=3).flatten(2)
    pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=3).flatten(2)
    return torch.cat((pos_y, pos_x), dim=2).reshape(h * w, 2 * num_pos_feats)


def _ln(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def multi_head_attention(p, prefix, query, key, value, num_heads=8):
    """layers.py:236-311 with the fused in_proj_weight [C, 3C] sliced per q/k/v (:221-234)."""
    C = query.shape[-1]
    D = C // num_heads
    Wi, bi = p[prefix + "in_proj_weight"], p[prefix + "in_proj_bias"]

    def proj(t, i):
        y = _store(t @ Wi[:, i * C:(i + 1) * C] + bi[i * C:(i + 1) * C])
        return y.reshape(y.shape[0], y.shape[1], num_heads, D).permute(0, 2, 1, 3)
    q, k, v = proj(query, 0), proj(key, 1), proj(value, 2)
    prod = (q @ k.transpose(-1, -2)) * (float(D) ** -0.5)
    w = F.softmax(prod, dim=-1)
    out = _store((w @ v).permute(0, 2, 1, 3).reshape(query.shape[0], query.shape[1], C))
    return out @ p[prefix + "out_proj.weight"] + p[prefix + "out_proj.bias"]


def _sub(p, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix)}


def encoder_layer_forward(p, prefix, src, ref, shapes, mask, pos):
    """TransformerEncoderLayer.forward (transformer_encoder_decoder.py:184-204), eval mode (dropout off)."""
    bs, _, c = src.shape
    start, _ = level_tables(shapes)
    branch = []
    for l, (h, w) in enumerate(shapes):                                                   # :163-196
        x = src[:, start[l]:start[l] + h * w].permute(0, 2, 1).reshape(bs, c, h, w)
        y = _store(F.conv2d(x, p[f"{prefix}conv{l}.0.weight"], None, 1, 1))
        y = F.group_norm(y, 32, p[f"{prefix}conv{l}.1.weight"], p[f"{prefix}conv{l}.1.bias"], 1e-5)
        y = F.gelu(y) + x
        branch.append(y.flatten(2).permute(0, 2, 1))
    src_flatten = torch.cat(branch, 1)
    if _ROUND["on"]:        # with_pos_embed folded into the query projection (row-bias table), as the kernels do
        src2 = msda_forward(_sub(p, prefix + "self_attn."), src, ref, src, shapes, mask, dtype=src.dtype, query_pos=pos[:1])
    else:
        src2 = msda_forward(_sub(p, prefix + "self_attn."), src + pos, ref, src, shapes, mask, dtype=src.dtype)  # :198
    src = _store(_ln(src + src2, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"]))    # :199-200
    # (the fused FFN kernel, emrt_ffn_fused_fwd: the hidden activations are rounded to bf16 on their way from the first
    # GEMM's accumulator to the second GEMM's operand — they never reach HBM, but the operand format is bf16 — and
    # x + linear2(...) is rounded to bf16 once when it is parked for the LayerNorm; norm2 and the final add follow in fp32)
    ffn = _store(F.relu(src @ p[prefix + "linear1.weight"] + p[prefix + "linear1.bias"])) @ p[prefix + "linear2.weight"] \
        + p[prefix + "linear2.bias"]                                                        # :157-158
    src = _ln(_store(src + ffn), p[prefix + "norm2.weight"], p[prefix + "norm2.bias"])     # :159-160
    return _store(src + src_flatten)                                                        # :203


def decoder_layer_forward(p, prefix, tgt, ref, memory, shapes, mask, query_pos):
    """TransformerDecoderLayer.forward (transformer_encoder_decoder.py:282-295), eval mode."""
    q = tgt + query_pos
    tgt2 = multi_head_attention(p, prefix + "self_attn.", q, q, tgt)
    tgt = _store(_ln(tgt + tgt2, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"]))
    if _ROUND["on"]:
        tgt2 = msda_forward(_sub(p, prefix + "cross_attn."), tgt, ref, memory, shapes, mask, dtype=tgt.dtype, query_pos=query_pos[:1])
    else:
        tgt2 = msda_forward(_sub(p, prefix + "cross_attn."), tgt + query_pos, ref, memory, shapes, mask, dtype=tgt.dtype)
    tgt = _store(_ln(tgt + tgt2, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"]))
    ffn = _store(F.relu(tgt @ p[prefix + "linear1.weight"] + p[prefix + "linear1.bias"])) @ p[prefix + "linear2.weight"] \
        + p[prefix + "linear2.bias"]
    return _store(_ln(tgt + ffn, p[prefix + "norm3.weight"], p[prefix + "norm3.bias"]))


def masked_position_embedding_sine(mask, num_pos_feats=128, temperature=10000.0, offset=-0.5, eps=1e-6, scale=2 * math.pi):
    """PositionEmbedding.forward (position_encoding.py:51-75) for a [B, h, w] 0/1 mask -> [B, h*w, 2*num_pos_feats]."""
    dtype = mask.dtype
    y_embed = mask.cumsum(1)
    x_embed = mask.cumsum(2)
    y_embed = (y_embed + offset) / (y_embed[:, -1:, :] + eps) * scale
    x_embed = (x_embed + offset) / (x_embed[:, :, -1:] + eps) * scale
    dim_t = 2 * (torch.arange(num_pos_feats) // 2).to(dtype)
    dim_t = temperature ** (dim_t / num_pos_feats)
    pos_x = x_embed[..., None] / dim_t
    pos_y = y_embed[..., None] / dim_t
    pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).flatten(1, 2)


def encoder_decoder_forward(p, src_feats, src_psp, num_enc=4, num_dec=2, trace=None, src_mask=None):
    """EncoderDecoder.forward (transformer_encoder_decoder.py:416-473), src_mask=None, eval mode.
    p: dict of torch tensors with Paddle state-dict keys (Linear [in,out], conv [out,in,kh,kw]).
    src_feats: [c2,c3,c4] NCHW; src_psp [B,256,110].  -> (hs [1,B,110,256], memory [B,Lv,256]).
    trace: optional dict, filled with the tensors between the layers (src, pos, enc[i], ref_dec, query_pos, dec[i]) — the
    per-layer (teacher-forced) parity tests feed each layer the evaluation's own input."""
    if src_mask is not None:
        return _encoder_decoder_forward_masked(p, src_feats, src_psp, num_enc, num_dec, src_mask)
    srcs, shapes = [], []
    for i, f in enumerate(src_feats):                                                       # :417-419
        y = _store(F.conv2d(f, p[f"input_proj.{i}.0.weight"], p[f"input_proj.{i}.0.bias"]))
        y = _store(F.group_norm(y, 32, p[f"input_proj.{i}.1.weight"], p[f"input_proj.{i}.1.bias"], 1e-5))
        srcs.append(y)
    src_flatten, pos_flatten = [], []
    for level, s in enumerate(srcs):                                                        # :434-452
        bs, c, h, w = s.shape
        shapes.append((h, w))
        src_flatten.append(s.flatten(2).permute(0, 2, 1))
        pos = position_embedding_sine(h, w, c // 2, dtype=s.dtype)
        pos_flatten.append((pos + p["level_embed.weight"][level].reshape(1, -1))[None].expand(bs, -1, -1))
    src = torch.cat(src_flatten, 1)
    pos = _store(torch.cat(pos_flatten, 1))
    bs = src.shape[0]
    mask = torch.ones(bs, src.shape[1], dtype=src.dtype)                                    # :451
    ref = encoder_reference_points(shapes, bs, src.dtype)
    out = src
    if trace is not None:
        trace.update(src=src, pos=pos, enc=[], dec=[])
    for i in range(num_enc):                                                                # :230-239
        out = encoder_layer_forward(p, f"encoder.layers.{i}.", out, ref, shapes, mask, pos)
        if trace is not None:
            trace["enc"].append(out)
    memory = out
    query_embed = p["query_pos_embed.weight"][None].expand(bs, -1, -1)                      # :464
    rp = torch.sigmoid(query_embed @ p["reference_points.weight"] + p["reference_points.bias"])  # :466
    rp = rp[:, :, None, :].expand(-1, -1, len(shapes), -1)                                  # :467 (valid_ratios == 1)
    tgt = src_psp.permute(0, 2, 1)                                                          # :469
    if trace is not None:
        trace.update(ref_dec=rp, query_pos=_store(query_embed), tgt=tgt)
    for i in range(num_dec):
        tgt = decoder_layer_forward(p, f"decoder.layers.{i}.", tgt, rp, memory, shapes, mask, _store(query_embed))
        if trace is not None:
            trace["dec"].append(tgt)
    return tgt[None], memory, shapes


def _encoder_decoder_forward_masked(p, src_feats, src_psp, num_enc, num_dec, src_mask):
    """EncoderDecoder.forward with src_mask [B, H, W] (non-zero = valid), transformer_encoder_decoder.py:408-473."""
    dt = src_psp.dtype
    srcs = []
    for i, f in enumerate(src_feats):                                                       # :417-419
        y = F.conv2d(f, p[f"input_proj.{i}.0.weight"], p[f"input_proj.{i}.0.bias"])
        srcs.append(F.group_norm(y, 32, p[f"input_proj.{i}.1.weight"], p[f"input_proj.{i}.1.bias"], 1e-5))
    src_flatten, mask_flatten, pos_flatten, shapes, valid_ratios = [], [], [], [], []
    sm = torch.as_tensor(src_mask).to(dt)
    for level, s in enumerate(srcs):                                                        # :434-452
        bs, c, h, w = s.shape
        shapes.append((h, w))
        src_flatten.append(s.flatten(2).permute(0, 2, 1))
        mask = (F.interpolate(sm[None], size=(h, w))[0] != 0).to(dt)                         # :440 (nearest), bool
        vr_h = mask[:, :, 0].sum(1) / h                                                      # :408-415
        vr_w = mask[:, 0, :].sum(1) / w
        valid_ratios.append(torch.stack([vr_w, vr_h], -1))
        pos = masked_position_embedding_sine(mask, c // 2)                                   # :446
        pos_flatten.append(pos + p["level_embed.weight"][level].reshape(1, 1, -1))           # :447
        mask_flatten.append(mask.flatten(1))
    src, pos, mask = torch.cat(src_flatten, 1), torch.cat(pos_flatten, 1), torch.cat(mask_flatten, 1)
    vr = torch.stack(valid_ratios, 1)                                                       # [bs, L, 2]
    bs = src.shape[0]
    # TransformerEncoder.get_reference_points (:213-228)
    pts = []
    for i, (H, W) in enumerate(shapes):
        ry, rx = torch.meshgrid(torch.linspace(0.5, H - 0.5, H, dtype=dt), torch.linspace(0.5, W - 0.5, W, dtype=dt), indexing="ij")
        ry = ry.flatten()[None] / (vr[:, None, i, 1] * H)
        rx = rx.flatten()[None] / (vr[:, None, i, 0] * W)
        pts.append(torch.stack((rx, ry), -1))
    ref = torch.cat(pts, 1)[:, :, None] * vr[:, None]                                       # [bs, Lv, L, 2]
    out = src
    for i in range(num_enc):
        out = encoder_layer_forward(p, f"encoder.layers.{i}.", out, ref, shapes, mask, pos)
    memory = out
    query_embed = p["query_pos_embed.weight"][None].expand(bs, -1, -1)
    rp = torch.sigmoid(query_embed @ p["reference_points.weight"] + p["reference_points.bias"])
    rp = rp[:, :, None, :] * vr[:, None]                                                    # :467
    tgt = src_psp.permute(0, 2, 1)
    for i in range(num_dec):
        tgt = decoder_layer_forward(p, f"decoder.layers.{i}.", tgt, rp, memory, shapes, mask, query_embed)
    return tgt[None], memory, shapes


def make_encoder_decoder_params(seed=1234, C=256, ffn=1024, heads=8, levels=3, points=6, num_enc=4, num_dec=2,
                                in_channels=(512, 1024, 2048), num_queries=110) -> Dict[str, torch.Tensor]:
    """Synthetic EncoderDecoder weights (SURVEY.md §8d distributions), Paddle key names and layouts."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p: Dict[str, np.ndarray] = {}
    xav = lambda i, o: math.sqrt(6.0 / (i + o))

    def linear(name, i, o):
        p[name + ".weight"] = rng_uniform(rng, (i, o), xav(i, o))
        p[name + ".bias"] = rng_uniform(rng, (o,), 0.1)

    def norm(name, c):
        p[name + ".weight"] = rng.uniform(0.5, 1.5, size=(c,)).astype(np.float32)
        p[name + ".bias"] = rng_normal(rng, (c,), 0.1)

    def msda(name):
        sub = make_msda_params(int(rng.integers(1 << 30)), C, heads, levels, points)
        for k, v in sub.items():
            p[name + "." + k] = v

    for i in range(num_enc):
        pre = f"encoder.layers.{i}"
        msda(pre + ".self_attn")
        norm(pre + ".norm1", C)
        linear(pre + ".linear1", C, ffn)
        linear(pre + ".linear2", ffn, C)
        norm(pre + ".norm2", C)
        for l in range(levels):
            p[f"{pre}.conv{l}.0.weight"] = rng_uniform(rng, (C, C, 3, 3), xav(C * 9, C * 9))
            norm(f"{pre}.conv{l}.1", C)
    for i in range(num_dec):
        pre = f"decoder.layers.{i}"
        p[pre + ".self_attn.in_proj_weight"] = rng_uniform(rng, (C, 3 * C), xav(C, 3 * C))
        p[pre + ".self_attn.in_proj_bias"] = rng_uniform(rng, (3 * C,), 0.1)
        linear(pre + ".self_attn.out_proj", C, C)
        norm(pre + ".norm1", C)
        msda(pre + ".cross_attn")
        norm(pre + ".norm2", C)
        linear(pre + ".linear1", C, ffn)
        linear(pre + ".linear2", ffn, C)
        norm(pre + ".norm3", C)
    p["level_embed.weight"] = rng_normal(rng, (levels, C))
    p["tgt_embed.weight"] = rng_normal(rng, (num_queries, C))
    p["query_pos_embed.weight"] = rng_normal(rng, (num_queries, C))
    linear("reference_points", C, 2)
    for i, cin in enumerate(in_channels):
        p[f"input_proj.{i}.0.weight"] = rng_uniform(rng, (C, cin, 1, 1), xav(cin, C))
        p[f"input_proj.{i}.0.bias"] = rng_uniform(rng, (C,), 0.1)
        norm(f"input_proj.{i}.1", C)
    return {k: torch.from_numpy(v) for k, v in p.items()}
