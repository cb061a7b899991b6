"""GPU parity tests of the head tail: upsample, window accumulation, finalize, fused stitch+argmax, areas."""
import os

import numpy as np
import pytest
import torch

import oracle as O
import emrt_b200
from emrt_b200 import ops, infer

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_upsample2x_matches_oracle(cuda_dev):
    rng = np.random.Generator(np.random.PCG64(0))
    for shape in [(2, 6, 16, 20), (1, 7, 5, 3), (3, 1, 1, 1)]:
        x = O.rng_normal(rng, shape)
        got = ops.upsample2x(torch.from_numpy(x).to(cuda_dev)).cpu()
        assert (got - O.upsample2x(x)).abs().max() < 1e-5
        assert np.abs(got.numpy() - O.upsample2x_loop(x)).max() < 1e-5
        xb = torch.from_numpy(x).bfloat16()
        gotb = ops.upsample2x(xb.to(cuda_dev)).cpu()
        assert (gotb - O.upsample2x(xb.float())).abs().max() < 1e-5


class LinearPixelModel:
    """logits = W @ rgb per pixel; exposes the half-resolution logits the fused path consumes."""

    def __init__(self, wmat, dev):
        self.w = torch.from_numpy(wmat).to(dev)

    def __call__(self, batch):
        return (torch.einsum("oc,nchw->nohw", self.w, batch.float()),)


def test_slide_inference_golden(cuda_dev):
    g = np.load(os.path.join(GOLD, "slide_small.npz"))
    model = LinearPixelModel(g["wmat"], cuda_dev)
    imgs = [torch.from_numpy(g["img0"]).to(cuda_dev), torch.from_numpy(g["img1"]).to(cuda_dev)]
    crop, stride = tuple(int(v) for v in g["crop"]), tuple(int(v) for v in g["stride"])
    logits = emrt_b200.slide_inference(model, imgs, crop, stride, 6)
    for i in range(2):
        assert tuple(logits[i].shape) == tuple(g[f"logit{i}"].shape)
        assert np.abs(logits[i].cpu().numpy() - g[f"logit{i}"]).max() < 1e-4
    ori = [tuple(int(v) for v in o) for o in g["ori"]]
    preds = emrt_b200.ss_inference(model, imgs, ori, True, None, stride, crop, 6)
    for i in range(2):
        assert preds[i].dtype == torch.int32 and tuple(preds[i].shape) == (1, 1) + ori[i]
        assert (preds[i].cpu().numpy() == g[f"pred{i}"]).mean() >= 0.999


class HalfModel:
    """A model whose full-resolution logits are UpHead's x2 upsample of a half-resolution tensor."""

    def __init__(self, wmat, dev, dtype=torch.float32):
        self.w = torch.from_numpy(wmat).to(dev)
        self.dtype = dtype

    def forward_half_logits(self, batch):
        half = torch.nn.functional.avg_pool2d(batch.float(), 2)
        return torch.einsum("oc,nchw->nohw", self.w, half).to(self.dtype).contiguous()

    def __call__(self, batch):
        return (ops.upsample2x(self.forward_half_logits(batch)),)


@pytest.mark.parametrize("H,W,crop,stride,nc", [(1024, 1024, 512, 384, 7), (96, 130, 64, 48, 6), (200, 200, 64, 40, 6)])
def test_fused_stitch_argmax_matches_oracle_and_unfused(cuda_dev, H, W, crop, stride, nc):
    rng = np.random.Generator(np.random.PCG64(2))
    wmat = O.rng_normal(rng, (nc, 3), 0.8)
    imgs = [torch.from_numpy(O.rng_normal(rng, (3, H, W))).to(cuda_dev) for _ in range(2)]
    model = HalfModel(wmat, cuda_dev)
    ori = [(H, W)] * 2
    fused = emrt_b200.ss_inference(model, imgs, ori, True, None, (stride, stride), (crop, crop), nc)

    class NoHalf:           # same model without the fast-path hook -> canvas path
        def __call__(self, b):
            return model(b)
    unfused = emrt_b200.ss_inference(NoHalf(), imgs, ori, True, None, (stride, stride), (crop, crop), nc)

    def cpu_model(b):       # oracle: same half logits, oracle upsample
        half = torch.nn.functional.avg_pool2d(b, 2)
        return (O.upsample2x(torch.einsum("oc,nchw->nohw", torch.from_numpy(wmat), half)),)
    want = O.ss_inference(cpu_model, [i.cpu() for i in imgs], ori, True, None, (stride, stride), (crop, crop), nc)
    for i in range(2):
        assert (fused[i].cpu() == want[i]).float().mean().item() >= 0.999
        assert (unfused[i].cpu() == want[i]).float().mean().item() >= 0.999
        assert (fused[i] == unfused[i]).float().mean().item() >= 0.9999


def test_finalize_resize_softmax_argmax(cuda_dev):
    rng = np.random.Generator(np.random.PCG64(3))
    logit = O.rng_normal(rng, (1, 6, 37, 53))
    for shape in [(37, 53), (80, 64), (20, 100)]:
        labels, probs, _ = ops.finalize_argmax(torch.from_numpy(logit).to(cuda_dev), None, out_hw=shape, want_probs=True)
        want = O.ss_inference_tail(logit, shape)
        assert (labels.cpu() == want).float().mean().item() >= 0.999
        wp = torch.softmax(O.interpolate_bilinear(logit, shape), 1)
        assert (probs.cpu() - wp).abs().max() < 1e-5
    lab8, _, _ = ops.finalize_argmax(torch.from_numpy(logit).to(cuda_dev), None, label_dtype=torch.uint8)
    assert lab8.dtype == torch.uint8 and torch.equal(lab8.cpu().int(), O.ss_inference_tail(logit, (37, 53)))
    # ties -> first maximal index
    tie = torch.zeros(1, 6, 4, 4)
    tie[:, 2] = 1.0
    tie[:, 4] = 1.0
    lab, _, _ = ops.finalize_argmax(tie.to(cuda_dev), None)
    assert torch.all(lab == 2)


def test_calculate_area_matches_oracle(cuda_dev):
    rng = np.random.Generator(np.random.PCG64(4))
    n = 512 * 512
    pred = rng.integers(0, 6, size=n).astype(np.int32)
    label = rng.integers(0, 6, size=n).astype(np.int32)
    label[rng.uniform(size=n) < 0.02] = 255
    got = emrt_b200.calculate_area(torch.from_numpy(pred).to(cuda_dev).reshape(1, 1, 512, 512),
                                   torch.from_numpy(label).to(cuda_dev).reshape(1, 1, 512, 512), 6).cpu().numpy()
    ia, pa, la = O.calculate_area(pred, label, 6)
    assert np.array_equal(got[0], ia) and np.array_equal(got[1], pa) and np.array_equal(got[2], la)
    with pytest.raises(ValueError):
        emrt_b200.calculate_area(torch.zeros(4, device=cuda_dev), torch.zeros(5, device=cuda_dev), 6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_quad_stitch_kernel_equals_pixel_kernel_any_window_parity(cuda_dev, dtype):
    """The 2x2-quad stitch kernel (3x3 patch per window) must equal the one-pixel kernel bit for bit — logits and
    labels — including windows at ODD origins (quads straddling window borders) and partially covered pixels."""
    import os
    rng = np.random.Generator(np.random.PCG64(21))
    nc, H, W, hc, wc = 7, 70, 90, 32, 48
    wins = [(0, 0, 0), (0, 0, 42), (0, 17, 5), (0, 38, 11), (0, 38, 42), (0, 3, 33), (1, 0, 0), (1, 38, 42), (1, 19, 21)]
    # fill the rest so every pixel is covered at least once
    for img in (0, 1):
        for y in range(0, H - hc + 1, 19):
            for x in range(0, W - wc + 1, 21):
                wins.append((img, y, x))
        wins += [(img, H - hc, x) for x in range(0, W - wc + 1, 21)] + [(img, y, W - wc) for y in range(0, H - hc + 1, 19)]
        wins.append((img, H - hc, W - wc))
    wins.sort(key=lambda w: w[0])
    half = torch.from_numpy(O.rng_normal(rng, (len(wins), nc, hc // 2, wc // 2))).to(cuda_dev).to(dtype)
    t = lambda k: torch.tensor([w[k] for w in wins], dtype=torch.int32, device=cuda_dev)
    lab_q, log_q = ops.stitch_argmax_fused(half, t(0), t(1), t(2), 2, H, W, want_logits=True)
    os.environ["EMRT_STITCH_PIXEL"] = "1"
    try:
        lab_p, log_p = ops.stitch_argmax_fused(half, t(0), t(1), t(2), 2, H, W, want_logits=True)
    finally:
        del os.environ["EMRT_STITCH_PIXEL"]
    assert torch.equal(log_q, log_p)
    assert torch.equal(lab_q, lab_p)
    # and against the oracle's canvas formulation
    full = O.upsample2x(half.float().cpu())
    canvas, cnt = torch.zeros(2, nc, H, W), torch.zeros(2, 1, H, W)
    for k, (i, y, x) in enumerate(wins):
        canvas[i, :, y:y + hc, x:x + wc] += full[k]
        cnt[i, :, y:y + hc, x:x + wc] += 1
    assert (cnt > 0).all()
    assert torch.allclose(log_q.cpu(), canvas / cnt, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("dtype,gt_dtype", [(torch.float32, torch.int64), (torch.bfloat16, torch.uint8)])
def test_fused_eval_areas_and_palette_match_unfused(cuda_dev, dtype, gt_dtype):
    """argmax -> per-image calculate_area histogram (ignore-index contract) and palette image inside
    the stitch kernel == ss_inference followed by calculate_area (oracle) and a palette lookup, bit for bit."""
    rng = np.random.Generator(np.random.PCG64(31))
    nc, H, W, crop, stride = 7, 200, 264, 64, 40          # W not a multiple of the 32-pixel CTA tile
    wmat = O.rng_normal(rng, (nc, 3), 0.8)
    imgs = [torch.from_numpy(O.rng_normal(rng, (3, H, W))).to(cuda_dev) for _ in range(3)]
    gts = []
    for _ in imgs:
        g = rng.integers(0, nc, size=(1, H, W))
        g[rng.uniform(size=g.shape) < 0.05] = 255
        gts.append(torch.from_numpy(g).to(gt_dtype).to(cuda_dev))
    palette = torch.from_numpy(rng.integers(0, 256, size=(nc, 3)).astype(np.uint8))
    model = HalfModel(wmat, cuda_dev, dtype)
    preds, areas, color = emrt_b200.ss_inference_eval(model, imgs, gts, (stride, stride), (crop, crop), nc, palette=palette)
    want = emrt_b200.ss_inference(model, imgs, [(H, W)] * 3, True, None, (stride, stride), (crop, crop), nc)
    assert areas.shape == (3, 3, nc) and areas.dtype == torch.int64 and color.shape == (3, H, W, 3)
    for i in range(3):
        assert torch.equal(preds[i], want[i])
        ia, pa, la = O.calculate_area(want[i].cpu(), gts[i].cpu().long().reshape(1, 1, H, W), nc)
        got = areas[i].cpu().numpy()
        assert np.array_equal(got[0], np.asarray(ia, dtype=np.int64).reshape(-1))
        assert np.array_equal(got[1], np.asarray(pa, dtype=np.int64).reshape(-1))
        assert np.array_equal(got[2], np.asarray(la, dtype=np.int64).reshape(-1))
        assert torch.equal(color[i].cpu(), palette[want[i].cpu().reshape(H, W).long()])
    # areas only / palette only
    _, a2, c2 = emrt_b200.ss_inference_eval(model, imgs, gts, (stride, stride), (crop, crop), nc)
    assert c2 is None and torch.equal(a2, areas)
    _, a3, c3 = emrt_b200.ss_inference_eval(model, imgs, None, (stride, stride), (crop, crop), nc, palette=palette)
    assert a3 is None and torch.equal(c3, color)


@pytest.mark.parametrize("label_dtype", [torch.uint8, torch.int32])
@pytest.mark.parametrize("case", ["cfg3", "mixed", "uncovered"])
def test_strip_stitch_kernel_equals_quad_and_pixel_kernels(cuda_dev, label_dtype, case):
    """The 16x2-strip stitch kernel (label map only, bf16 logits) must give the SAME labels, bit for bit, as the quad and
    the one-pixel kernels: cfg-3 plan (origins {0, 384, 512}: all 16-aligned -> vector path), a plan mixing aligned and
    unaligned / odd origins (per-pixel path inside the strip kernel, strips cut by window borders), and a plan that
    leaves pixels uncovered (sum 0 for every class -> label 0)."""
    rng = np.random.Generator(np.random.PCG64(77))
    nc = 7
    if case == "cfg3":
        n_img, H, W, hc, wc = 2, 1024, 1024, 512, 512
        plan, _, _ = emrt_b200.plan_windows([(H, W)] * n_img, (wc, hc), (384, 384))
        wins = [(p[0], p[1], p[2]) for p in plan]
    else:
        n_img, H, W, hc, wc = 2, 96, 160, 32, 48
        wins = [(0, 0, 0), (0, 0, 48), (0, 16, 112), (0, 17, 5), (0, 38, 11), (0, 64, 96), (0, 3, 33), (1, 0, 0), (1, 64, 112), (1, 19, 21),
                (1, 32, 16), (0, 32, 32)]
        if case == "mixed":
            for img in range(n_img):
                for y in list(range(0, H - hc + 1, 16)) + [H - hc]:
                    for x in list(range(0, W - wc + 1, 32)) + [W - wc]:
                        wins.append((img, y, x))
        wins.sort(key=lambda w: w[0])
    half = torch.from_numpy(O.rng_normal(rng, (len(wins), nc, hc // 2, wc // 2))).to(cuda_dev).bfloat16()
    # ties: make a few windows' classes exactly equal so the first-max rule is exercised
    half[0, 1] = half[0, 0]
    half[-1, 3] = half[-1, 2]
    t = lambda k: torch.tensor([w[k] for w in wins], dtype=torch.int32, device=cuda_dev)
    run = lambda: ops.stitch_argmax_fused(half, t(0), t(1), t(2), n_img, H, W, label_dtype=label_dtype)[0]
    lab_s = run()
    outs = {}
    for env in ("EMRT_STITCH_QUAD", "EMRT_STITCH_PIXEL"):
        os.environ[env] = "1"
        try:
            outs[env] = run()
        finally:
            del os.environ[env]
    assert torch.equal(lab_s, outs["EMRT_STITCH_QUAD"]) and torch.equal(lab_s, outs["EMRT_STITCH_PIXEL"])
    assert lab_s.dtype == label_dtype and int(lab_s.max()) < nc
    if case == "uncovered":
        cover = torch.zeros(n_img, H, W)
        for (i, y, x) in wins:
            cover[i, y:y + hc, x:x + wc] += 1
        assert (cover == 0).any() and bool((lab_s.cpu()[:, 0][cover == 0] == 0).all())
