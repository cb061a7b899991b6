"""GPU parity of the decoder glue and of the whole EncoderDecoder against the oracle's
restatement of layers.py:144-311 and transformer_encoder_decoder.py:242-473."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle as O
import emrt_b200
from emrt_b200 import ops, _lib as L
from parity import assert_bf16_parity, assert_layers_match, oracle_encdec_pair

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    want = torch.as_tensor(want).double()
    return ((got.detach().double().cpu() - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


def _load(module, params, prefix=""):
    with torch.no_grad():
        sd = module.state_dict()
        for k in sd:
            sd[k].copy_(torch.as_tensor(params[prefix + k]))
    return module


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("Lq,Lk", [(110, 110), (37, 200)])
def test_mha_small_core(cuda_dev, dtype, Lq, Lk):
    B, M, D = 3, 8, 32
    rng = np.random.Generator(np.random.PCG64(Lq))
    qk = torch.from_numpy(O.rng_normal(rng, (B, max(Lq, Lk), 2 * M * D))).to(dtype)      # fused [q | k] projection buffer
    v = torch.from_numpy(O.rng_normal(rng, (B, Lk, M * D))).to(dtype)
    q, k = qk[:, :Lq, :M * D], qk[:, :Lk, M * D:]
    heads = lambda t: t.double().reshape(t.shape[0], t.shape[1], M, D).permute(0, 2, 1, 3)
    w = F.softmax(heads(q) @ heads(k).transpose(-1, -2) * D ** -0.5, -1)
    want = (w @ heads(v)).permute(0, 2, 1, 3).reshape(B, Lq, M * D)
    qkd = qk.to(cuda_dev)
    qd, kd = qkd[:, :Lq, :M * D], qkd[:, :Lk, M * D:]
    if Lq != Lk:       # row strides must be uniform per tensor: give q / k their own buffers
        qd, kd = qd.contiguous(), kd.contiguous()
    got = ops.mha_small(qd, kd, v.to(cuda_dev), M, D ** -0.5)
    assert rel_err(got.float(), want) < (1e-5 if dtype == torch.float32 else 1e-2)


def test_decoder_layer_fp32_matches_oracle(cuda_dev):
    shapes = [(16, 16), (8, 8), (4, 4)]
    B, C, Nq = 2, 256, 110
    params = O.make_encoder_decoder_params(31, num_enc=0, num_dec=1)
    rng = np.random.Generator(np.random.PCG64(32))
    _, Lv = O.level_tables(shapes)
    tgt = torch.from_numpy(O.rng_normal(rng, (B, Nq, C)))
    mem = torch.from_numpy(O.rng_normal(rng, (B, Lv, C)))
    qpos = torch.from_numpy(O.rng_normal(rng, (B, Nq, C), 0.5))
    ref = torch.from_numpy(np.repeat(rng.uniform(0.1, 0.9, size=(B, Nq, 1, 2)).astype(np.float32), 3, axis=2))
    p64 = {k: torch.as_tensor(v).double() for k, v in params.items()}
    want = O.decoder_layer_forward(p64, "decoder.layers.0.", tgt.double(), ref.double(), mem.double(), shapes,
                                   torch.ones(B, Lv).double(), qpos.double())
    layer = _load(emrt_b200.TransformerDecoderLayer(C, 8, 1024, 0.1, "relu", 3, 6), params, "decoder.layers.0.").to(cuda_dev)
    d = lambda t: t.to(cuda_dev)
    got = layer(d(tgt), d(ref), d(mem), torch.tensor(shapes), d(torch.ones(B, Lv)), d(qpos))
    assert rel_err(got, want) < 2e-4


def _features(rng, B, tile, dtype=torch.float32):
    chans = (512, 1024, 2048)
    feats = [torch.from_numpy(O.rng_normal(rng, (B, c, tile // s, tile // s), 0.5)).to(dtype) for c, s in zip(chans, (8, 16, 32))]
    psp = torch.from_numpy(O.rng_normal(rng, (B, 256, 110), 0.5)).to(dtype)
    return feats, psp


def test_encoder_decoder_fp32_matches_oracle(cuda_dev):
    """The whole EncoderDecoder.forward (input_proj + position / level embedding + 2 encoder layers + reference-point
    head + 2 decoder layers) in fp32 on a 128x128 tile's features."""
    params = O.make_encoder_decoder_params(41, num_enc=2, num_dec=2)
    rng = np.random.Generator(np.random.PCG64(42))
    feats, psp = _features(rng, 2, 128)
    p64 = {k: torch.as_tensor(v).double() for k, v in params.items()}
    whs, wmem, _ = O.encoder_decoder_forward(p64, [f.double() for f in feats], psp.double(), num_enc=2, num_dec=2)
    m = emrt_b200.EncoderDecoder(110, "sine", False, (512, 1024, 2048), 3, 6, 6, 6, 256, 8, 2, 2, 1024)
    m = _load(m, params).to(cuda_dev)
    hs, mem = m([f.to(cuda_dev) for f in feats], psp.to(cuda_dev))
    assert tuple(hs.shape) == (1, 2, 110, 256) and tuple(mem.shape) == tuple(wmem.shape)
    assert rel_err(mem, wmem) < 5e-4
    assert rel_err(hs, whs) < 5e-4


def test_encoder_decoder_bf16_matches_oracle(cuda_dev):
    """Same, bf16 activations on the B200 path (tcgen05 GEMMs / conv, window-staged gather) at a 256x256 tile, EMRT's
    real depth (4 encoder + 2 decoder layers), against the float64 oracle on the bf16-rounded inputs and matrices."""
    params = O.make_encoder_decoder_params(43, num_enc=4, num_dec=2)
    rng = np.random.Generator(np.random.PCG64(44))
    feats, psp = _features(rng, 2, 256, torch.bfloat16)
    trace = {}
    (whs, wmem), (rhs, rmem) = oracle_encdec_pair(params, feats, psp, 4, 2, trace=trace)
    m = emrt_b200.EncoderDecoder(110, "sine", False, (512, 1024, 2048), 3, 6, 6, 6, 256, 8, 4, 2, 1024)
    m = _load(m, params).to(cuda_dev)
    hs, mem = m([f.to(cuda_dev) for f in feats], psp.to(cuda_dev))
    assert hs.dtype == torch.bfloat16
    # six layers deep with every intermediate stored in bf16 / fp16 (tests/parity.py): each layer on the same-rounding
    # oracle's own input within 1e-3 (the kernels' own error), and end to end no further from the exact evaluation than
    # the storage formats alone are
    assert_layers_match(m, feats, trace, cuda_dev)
    assert_bf16_parity(mem.float(), wmem, rmem, "memory")
    assert_bf16_parity(hs.float(), whs, rhs, "hs")
    assert rel_err(mem.float(), wmem) < 1e-1 and rel_err(hs.float(), whs) < 1e-1


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,C,P", [(2, 512, 4096), (3, 256, 110), (1, 2048, 64), (2, 64, 192), (1, 96, 64)])
def test_nchw_to_tokens_is_an_exact_transpose(cuda_dev, dtype, B, C, P):
    """Both kernels (generic 32x32 tiles; bf16 64x64 tiles with 16-byte accesses when C and P are multiples of 64)."""
    x = torch.randn(B, C, P, generator=torch.Generator().manual_seed(P)).to(dtype)
    got = ops.nchw_to_tokens(x.to(cuda_dev))
    assert torch.equal(got.cpu(), x.permute(0, 2, 1).contiguous())
