"""Edge cases and error behaviour of the drop-ins (SURVEY.md 8b "Errors" row; the cases the reference's asserts /
TypeError / ValueError cover, plus the degenerate inputs a parity suite should not skip)."""
import numpy as np
import pytest
import torch

import oracle.emrt_oracle as O
import emrt_b200
from emrt_b200 import ops, _lib as L

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    want = torch.as_tensor(want).double()
    return ((got.detach().double().cpu() - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


def _module(dev, seed=3, L_=3):
    params = O.make_msda_params(seed, 256, 8, L_, 6)
    m = emrt_b200.MSDeformableAttention(256, 8, L_, 6).to(dev)
    with torch.no_grad():
        for name, arr in params.items():
            mod, leaf = name.split(".")
            getattr(getattr(m, mod), leaf).copy_(torch.from_numpy(arr))
    return m.requires_grad_(False), params


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_all_zero_value_mask_leaves_only_the_output_bias(cuda_dev, dtype):
    """value *= mask (t_e_d.py:84-86): with every value row masked the gather returns zeros and the module's output is
    exactly output_proj.bias."""
    m, params = _module(cuda_dev)
    shapes = [(8, 8), (4, 4), (2, 2)]
    rng = np.random.Generator(np.random.PCG64(0))
    q = torch.from_numpy(O.rng_normal(rng, (2, 84, 256))).to(cuda_dev).to(dtype)
    v = torch.from_numpy(O.rng_normal(rng, (2, 84, 256))).to(cuda_dev).to(dtype)
    ref = emrt_b200.get_reference_points(shapes, device=cuda_dev)
    out = m(q, ref, v, shapes, torch.zeros(2, 84, device=cuda_dev))
    want = torch.from_numpy(params["output_proj.bias"]).to(dtype).float()
    assert torch.equal(out.float().cpu(), want.expand(2, 84, 256))


def test_single_query_single_batch_and_far_away_samples(cuda_dev):
    """Lq = 1, B = 1; reference points far outside [0, 1] (every sample off the map -> exactly output_proj.bias)."""
    m, params = _module(cuda_dev)
    shapes = [(8, 8), (4, 4), (2, 2)]
    rng = np.random.Generator(np.random.PCG64(1))
    q = O.rng_normal(rng, (1, 1, 256))
    v = O.rng_normal(rng, (1, 84, 256))
    ref = rng.uniform(0.2, 0.8, size=(1, 1, 3, 2)).astype(np.float32)
    d = lambda a: torch.from_numpy(a).to(cuda_dev)
    out = m(d(q), d(ref), d(v), shapes)
    want = O.msda_forward(params, q, ref, v, shapes, None, 8, 6)
    assert rel_err(out, want) < 1e-4
    far = m(d(q), d(ref) + 50.0, d(v), shapes)
    assert torch.equal(far.cpu().reshape(-1), torch.from_numpy(params["output_proj.bias"]))
    far16 = m(d(q).bfloat16(), d(ref) + 50.0, d(v).bfloat16(), shapes)
    assert torch.equal(far16.float().cpu().reshape(-1), torch.from_numpy(params["output_proj.bias"]).bfloat16().float())


def test_nan_sampling_location_contributes_zero_not_garbage(cuda_dev):
    """A NaN location has no corner inside the map: weight zero (the oracle's grid_sample gives NaN-free zeros padding
    only for finite inputs; the kernels define the NaN case as 'no contribution' and must not read out of bounds)."""
    shapes = [(32, 32), (16, 16), (8, 8)]
    rng = np.random.Generator(np.random.PCG64(2))
    _, Lv = O.level_tables(shapes)
    value = torch.from_numpy(O.rng_normal(rng, (1, Lv, 8, 32))).to(cuda_dev)
    loc = torch.from_numpy(rng.uniform(0, 1, size=(1, Lv, 8, 3, 6, 2)).astype(np.float32)).to(cuda_dev)
    attn = torch.full((1, Lv, 8, 3, 6), 1.0 / 18, device=cuda_dev)
    clean = emrt_b200.deformable_attention_core_func(value, torch.tensor(shapes), loc, attn)
    loc_nan = loc.clone()
    loc_nan[0, 5, 3, 1, 2, 0] = float("nan")
    attn0 = attn.clone()
    attn0[0, 5, 3, 1, 2] = 0.0
    want = emrt_b200.deformable_attention_core_func(value, torch.tensor(shapes), loc, attn0)
    got = emrt_b200.deformable_attention_core_func(value, torch.tensor(shapes), loc_nan, attn)
    assert torch.isfinite(got).all() and torch.equal(got, want) and not torch.equal(got, clean)
    # bf16 kernels (L1 path and window-staged path: Lq == Lv with head-major values and the pixel-grid hint)
    v16 = value.bfloat16()
    got16 = emrt_b200.deformable_attention_core_func(v16, torch.tensor(shapes), loc_nan, attn)
    want16 = emrt_b200.deformable_attention_core_func(v16, torch.tensor(shapes), loc, attn0)
    assert torch.isfinite(got16.float()).all() and torch.equal(got16, want16)
    v_hm = v16.permute(0, 2, 1, 3).contiguous()
    mode = L.LOC_NORMALIZED | L.VALUE_HEAD_MAJOR | L.QUERY_PIXEL_GRID
    got_w = ops.msda_gather_fwd(v_hm, loc_nan.half(), attn.half(), shapes, mode=mode)
    want_w = ops.msda_gather_fwd(v_hm, loc.half(), attn0.half(), shapes, mode=mode)
    assert torch.isfinite(got_w.float()).all() and torch.equal(got_w, want_w)


def test_reference_error_behaviour(cuda_dev):
    m, _ = _module(cuda_dev)
    shapes = [(8, 8), (4, 4), (2, 2)]
    q = torch.zeros(1, 84, 256, device=cuda_dev)
    ref = emrt_b200.get_reference_points(shapes, device=cuda_dev)
    with pytest.raises(AssertionError):                                    # t_e_d.py:81: sum(h*w) == Len_v
        m(q, ref, torch.zeros(1, 80, 256, device=cuda_dev), shapes)
    with pytest.raises(AssertionError):                                    # t_e_d.py:34: head-dim divisibility
        emrt_b200.MSDeformableAttention(250, 8, 3, 6)
    with pytest.raises(emrt_b200.EmrtError):                               # no CPU fallback
        m(q.cpu(), ref.cpu(), q.cpu(), shapes)
    with pytest.raises(emrt_b200.EmrtError):
        m(q.half(), ref, q.half(), shapes)                                 # fp16 activations are not a supported dtype
    model = lambda b: torch.zeros(b.shape[0], 6, b.shape[2], b.shape[3], device=b.device)   # not a Sequence
    img = [torch.zeros(3, 64, 64, device=cuda_dev)]
    with pytest.raises(TypeError, match="collections.abc.Sequence"):       # infer.py:115-118
        emrt_b200.ss_inference(model, img[0], None, False, None, (48, 48), (64, 64), 6)
    with pytest.raises(ValueError, match="batch_size should be set to 1"): # infer.py:122-124
        emrt_b200.ss_inference(model, img * 2, None, False, None, (48, 48), (64, 64), 6)
    with pytest.raises(TypeError, match="logits must be one of"):          # infer.py:126-129
        emrt_b200.ss_inference(model, [img[0][None]], None, False, None, (48, 48), (64, 64), 6)
    # library-level argument checks surface as EmrtError with the library's message
    with pytest.raises(emrt_b200.EmrtError):
        ops.stitch_argmax_fused(torch.zeros(1, 6, 5, 5, device=cuda_dev), torch.zeros(1, dtype=torch.int32, device=cuda_dev),
                                torch.zeros(1, dtype=torch.int32, device=cuda_dev),
                                torch.zeros(1, dtype=torch.int32, device=cuda_dev), 1, 10, 10,
                                labels=torch.zeros(1, 1, 10, 10, dtype=torch.float32, device=cuda_dev))


def test_slide_inference_images_smaller_than_the_crop_and_ragged_sizes(cuda_dev):
    """infer.py:52-59 with H < h_crop (the window shrinks to the image), and ragged image sizes in one call.  The two
    cases run separately: the reference itself cannot batch windows of different sizes (its concat at infer.py:64 raises,
    and so does the oracle); ours groups windows by size and is also run on the mixed list."""
    rng = np.random.Generator(np.random.PCG64(5))
    nc, crop, stride = 6, (32, 32), (20, 20)
    wmat = O.rng_normal(rng, (nc, 3), 0.7)
    small = [O.rng_normal(rng, (3, 20, 28))]
    ragged = [O.rng_normal(rng, (3, 50, 41)), O.rng_normal(rng, (3, 32, 32)), O.rng_normal(rng, (3, 33, 70))]

    def cpu_model(b):
        return (torch.einsum("oc,nchw->nohw", torch.from_numpy(wmat), b),)

    class Model:
        def __call__(self, b):
            return (torch.einsum("oc,nchw->nohw", torch.from_numpy(wmat).to(b.device), b.float()),)
    results = {}
    for name, imgs, ori in (("small", small, [(20, 28)]), ("ragged", ragged, [(64, 48), (32, 32), (33, 70)])):
        want = O.slide_inference(cpu_model, [torch.from_numpy(i) for i in imgs], crop, stride, nc)
        dev_imgs = [torch.from_numpy(i).to(cuda_dev) for i in imgs]
        got = emrt_b200.slide_inference(Model(), dev_imgs, crop, stride, nc)
        for g, w in zip(got, want):
            assert tuple(g.shape) == tuple(w.shape) and rel_err(g, w) < 1e-5
        preds = emrt_b200.ss_inference(Model(), dev_imgs, ori, True, None, stride, crop, nc)
        wantp = O.ss_inference(cpu_model, [torch.from_numpy(i) for i in imgs], ori, True, None, stride, crop, nc)
        for g, w in zip(preds, wantp):
            assert g.dtype == torch.int32 and tuple(g.shape) == tuple(w.shape)
            assert (g.cpu() == w).float().mean().item() >= 0.999
        results[name] = got
    with pytest.raises(RuntimeError):
        O.slide_inference(cpu_model, [torch.from_numpy(i) for i in small + ragged], crop, stride, nc)
    mixed = emrt_b200.slide_inference(Model(), [torch.from_numpy(i).to(cuda_dev) for i in small + ragged], crop, stride, nc)
    for g, w in zip(mixed, results["small"] + results["ragged"]):
        assert torch.equal(g, w)
