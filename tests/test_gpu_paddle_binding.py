"""Executes emrt_b200/paddle_shim.py — the binding a maintainer adds to the reference — on a
B200.  PaddlePaddle cannot be installed in this image, so the binding runs on oracle/paddle_on_torch.py, the torch-backed
stand-in for the Paddle API (device tensors, nn.Layer / nn.Linear with [in, out] weights, autograd.PyLayer, the current
CUDA stream): every ctypes call of the binding — argument order, dtypes, shapes, workspaces — reaches the real kernels
and is checked against the vectors generated from the reference's own code (tests/golden/ref_*.npz).  What this cannot
show is Paddle's own allocator / stream plumbing; it does show the binding's code is executable and correct."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import oracle.emrt_oracle as O

from parity import assert_bf16_parity, oracle_encdec_pair

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)
import make_reference_vectors as G  # noqa: E402

load = lambda name: np.load(os.path.join(GOLD, name + ".npz"))
raw = lambda t: t.detach().as_subclass(torch.Tensor)


def rel_err(got, want):
    want = torch.as_tensor(np.asarray(want)).double()
    return ((raw(got).double().cpu() - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


def l2_err(got, want):
    want = torch.as_tensor(np.asarray(want)).double()
    return ((raw(got).double().cpu() - want).norm() / want.norm()).item()


@pytest.fixture(scope="module")
def binding(cuda_dev):
    from oracle import paddle_on_torch as P
    paddle = P.install()
    paddle.set_device("gpu")
    import emrt_b200.paddle_shim as PS
    yield PS, paddle
    paddle.set_device("cpu")


def _module(PS, paddle, c):
    m = PS.MSDeformableAttention(c["C"], c["M"], len(c["shapes"]), c["P"])
    for name, arr in c["params"].items():
        mod, leaf = name.split(".")
        getattr(getattr(m, mod), leaf).set_value(paddle.to_tensor(arr))
    m.invalidate_packed()
    return m


def test_binding_msda_fp32_reference_composition_around_native_gather(binding):
    PS, paddle = binding
    g, c = load("ref_msda"), G.msda_inputs()
    m = _module(PS, paddle, c)
    T = paddle.to_tensor
    with paddle.no_grad():
        out = m(T(c["query"]), T(c["ref"]), T(c["value"]), T(c["shapes"], dtype="int64"), T(c["mask"]))
    assert out.shape == list(g["out"].shape) and rel_err(out, g["out"]) < 1e-4


def test_binding_msda_bf16_fused_inference_path(binding):
    PS, paddle = binding
    g, c = load("ref_msda"), G.msda_inputs()
    m = _module(PS, paddle, c)
    T = paddle.to_tensor
    with paddle.no_grad():
        out = m(T(c["query"]).astype("bfloat16"), T(c["ref"]), T(c["value"]).astype("bfloat16"), T(c["shapes"], dtype="int64"),
                T(c["mask"]))
    assert out.dtype == paddle.bfloat16
    # kernels' own error vs the same-rounding-points oracle; distance to the reference = input / weight rounding (tests/parity.py)
    r16 = lambda a: torch.as_tensor(a).bfloat16().double()
    p64 = {k: (r16(v) if k.endswith("weight") else torch.as_tensor(v).double()) for k, v in c["params"].items()}
    with O.kernel_storage_rounding():
        rounded = O.msda_forward(p64, r16(c["query"]), c["ref"], r16(c["value"]), c["shapes"], c["mask"], c["M"], c["P"],
                                 dtype=torch.float64).bfloat16().double()
    assert_bf16_parity(raw(out).float(), g["out"], rounded, "binding MSDA bf16 vs reference")
    # the encoder-shaped call (Lq == Lv, pixel-centre reference points) takes the window-staged gather
    shapes = [(32, 32), (16, 16), (8, 8)]
    rng = np.random.Generator(np.random.PCG64(5))
    _, Lv = O.level_tables(shapes)
    params = O.make_msda_params(5, 256, 8, 3, 6)
    q, v = O.rng_normal(rng, (2, Lv, 256)), O.rng_normal(rng, (2, Lv, 256))
    ref = O.encoder_reference_points(shapes, 2).numpy()
    r = lambda a: torch.from_numpy(a).bfloat16().float().numpy()
    want = O.msda_forward({k: (r(a) if k.endswith("weight") else a) for k, a in params.items()}, r(q), ref, r(v), shapes, None, 8, 6,
                          dtype=torch.float64)
    m2 = _module(PS, paddle, dict(C=256, M=8, shapes=shapes, P=6, params=params))
    with paddle.no_grad():
        out2 = m2(T(q).astype("bfloat16"), T(ref), T(v).astype("bfloat16"), T(shapes, dtype="int64"))
    assert rel_err(out2, want) < 1e-2


def test_binding_training_path_gradients(binding):
    """Gradients through the binding (Paddle autograd for the Linears / softmax, PyLayer for the native gather backward)
    against torch autograd through the float64 oracle."""
    PS, paddle = binding
    c = G.msda_inputs()
    m = _module(PS, paddle, c)
    T = paddle.to_tensor
    q, v = T(c["query"]), T(c["value"])
    q.stop_gradient = False
    v.stop_gradient = False
    out = m(q, T(c["ref"]), v, T(c["shapes"], dtype="int64"), T(c["mask"]))
    rng = np.random.Generator(np.random.PCG64(77))
    w = O.rng_normal(rng, tuple(out.shape))
    params = [p for _, p in sorted(m.named_parameters())]
    got = torch.autograd.grad((raw_keep(out) * torch.from_numpy(w).cuda()).sum(), [q, v] + params)
    # oracle
    tq = torch.from_numpy(c["query"]).double().requires_grad_()
    tv = torch.from_numpy(c["value"]).double().requires_grad_()
    tp = {k: torch.from_numpy(a).double().requires_grad_() for k, a in c["params"].items()}
    oo = O.msda_forward(tp, tq, c["ref"], tv, c["shapes"], c["mask"], c["M"], c["P"], dtype=torch.float64)
    want = torch.autograd.grad((oo * torch.from_numpy(w).double()).sum(), [tq, tv] + [tp[k] for k in sorted(tp)])
    for a, b in zip(got, want):
        assert rel_err(a, b) < 5e-4


def raw_keep(t):
    return t.as_subclass(torch.Tensor)


def test_binding_core_func_forward_and_backward(binding):
    PS, paddle = binding
    g, c = load("ref_core"), G.core_inputs()
    T = paddle.to_tensor
    value, loc, attn = T(c["value"]), T(c["loc"]), T(c["attn"])
    for t in (value, loc, attn):
        t.stop_gradient = False
    out = PS.deformable_attention_core_func(value, T(c["shapes"], dtype="int64"), loc, attn)
    assert rel_err(out, g["out"]) < 1e-4
    rng = np.random.Generator(np.random.PCG64(78))
    w = O.rng_normal(rng, tuple(out.shape))
    got = torch.autograd.grad((raw_keep(out) * torch.from_numpy(w).cuda()).sum(), [value, loc, attn])
    tv, tl, ta = (torch.from_numpy(c[k]).double().requires_grad_() for k in ("value", "loc", "attn"))
    oo = O.deformable_attention_core_func(tv, c["shapes"], tl, ta)
    want = torch.autograd.grad((oo * torch.from_numpy(w).double()).sum(), [tv, tl, ta])
    assert rel_err(got[0], want[0]) < 1e-4 and rel_err(got[2], want[2]) < 1e-4
    assert l2_err(got[1], want[1]) < 1e-3          # d/d loc is discontinuous at pixel borders: L2


def test_binding_slide_and_ss_inference(binding):
    PS, paddle = binding
    from oracle import paddle_on_torch as P
    from emrt_b200 import ops
    g, c = load("ref_slide"), G.slide_inputs()
    wconv = torch.from_numpy(c["wconv"]).cuda()

    def model(batch):
        half = torch.nn.functional.conv2d(raw(batch).float(), wconv, stride=2).contiguous()
        return (P._wrap(ops.upsample2x(half)),)
    imgs = [paddle.to_tensor(i) for i in c["imgs"]]
    logits = PS.slide_inference(model, imgs, c["crop"], c["stride"], c["nc"])
    for i in range(2):
        assert logits[i].shape == list(g[f"logit{i}"].shape) and rel_err(logits[i], g[f"logit{i}"]) < 1e-5
    preds = PS.ss_inference(model, imgs, c["ori"], True, None, c["stride"], c["crop"], c["nc"])
    for i in range(2):
        assert preds[i].dtype == paddle.int32 and preds[i].shape == list(g[f"pred{i}"].shape)
        assert (raw(preds[i]).cpu().numpy() == g[f"pred{i}"]).mean() >= 0.999


def test_binding_patch_reference_installs_the_drop_ins(binding):
    """patch_reference() rebinds the names the reference binds at import (t_e_d.py:14, val.py's `infer.ss_inference`)."""
    PS, paddle = binding
    names = ["src", "src.api", "src.api.infer", "src.models", "src.models.EMRT_utils",
             "src.models.EMRT_utils.transformer_encoder_decoder", "src.models.EMRT_utils.utils"]
    saved = {n: sys.modules.get(n) for n in names}
    try:
        for n in names:
            sys.modules[n] = types.ModuleType(n)
        for n in names[1:]:
            parent, leaf = n.rsplit(".", 1)
            setattr(sys.modules[parent], leaf, sys.modules[n])
        infer = sys.modules["src.api.infer"]
        infer.ss_inference = infer.slide_inference = object()
        original = infer.ss_inference
        PS.patch_reference()
        ted, U = sys.modules["src.models.EMRT_utils.transformer_encoder_decoder"], sys.modules["src.models.EMRT_utils.utils"]
        assert ted.MSDeformableAttention is PS.MSDeformableAttention
        assert ted.deformable_attention_core_func is PS.deformable_attention_core_func is U.deformable_attention_core_func
        assert infer.ss_inference is PS.ss_inference and infer.slide_inference is PS.slide_inference
        assert infer._emrt_original_ss_inference is original
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m


def test_binding_packs_follow_the_parameters(binding):
    """The bf16 weight packs are keyed on (data_ptr, inplace_version) of the parameters: a later set_value (checkpoint load,
    optimiser step) is seen by the next forward without any invalidate call."""
    PS, paddle = binding
    c = G.msda_inputs()
    m = PS.MSDeformableAttention(c["C"], c["M"], len(c["shapes"]), c["P"])
    for name, arr in c["params"].items():
        mod, leaf = name.split(".")
        getattr(getattr(m, mod), leaf).set_value(paddle.to_tensor(arr))
    T = paddle.to_tensor
    args = lambda: (T(c["query"]).astype("bfloat16"), T(c["ref"]), T(c["value"]).astype("bfloat16"), T(c["shapes"], dtype="int64"))
    with paddle.no_grad():
        a = raw(m(*args())).float().clone()
        m.output_proj.bias.set_value(paddle.to_tensor(c["params"]["output_proj.bias"] + 1.0))
        b = raw(m(*args())).float()
    assert (b - a - 1.0).abs().max().item() < 5e-2


def test_binding_patch_reference_is_idempotent_and_installs_half_logits(binding):
    """patch_reference() twice keeps the REFERENCE's ss_inference as the saved original (a second call used to save the
    shim itself -> infinite recursion for is_slide=False), and gives the model class `forward_half_logits`."""
    PS, paddle = binding
    names = ["src", "src.api", "src.api.infer", "src.models", "src.models.paddle_EMRT", "src.models.EMRT_utils",
             "src.models.EMRT_utils.transformer_encoder_decoder", "src.models.EMRT_utils.utils"]
    saved = {n: sys.modules.get(n) for n in names}
    try:
        for n in names:
            sys.modules[n] = types.ModuleType(n)
        for n in names[1:]:
            parent, leaf = n.rsplit(".", 1)
            setattr(sys.modules[parent], leaf, sys.modules[n])
        infer = sys.modules["src.api.infer"]
        original = infer.ss_inference = lambda *a, **k: "reference"
        infer.slide_inference = object()
        emrt_mod = sys.modules["src.models.paddle_EMRT"]

        class UpHead:
            num_conv, align_corners = 3, False
            def forward(self, x):
                return "full"

        class EMRT:
            pass
        emrt_mod.UpHead, emrt_mod.EMRT = UpHead, EMRT
        PS.patch_reference()
        PS.patch_reference()
        assert infer._emrt_original_ss_inference is original
        assert PS.ss_inference(None, [0], None, False, None, None, None, 6) == "reference"      # no recursion
        assert hasattr(EMRT, "forward_half_logits") and UpHead._emrt_wrapped
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m


def test_binding_ss_inference_takes_the_fused_kernel_with_half_logits(binding):
    """val.py:145's call on a model that exposes forward_half_logits (what install_half_logits gives the reference's EMRT):
    the shim's ss_inference runs emrt_stitch_argmax_fused on the half-resolution logits — same labels as the canvas path."""
    PS, paddle = binding
    from oracle import paddle_on_torch as P
    from emrt_b200 import ops
    g, c = load("ref_slide"), G.slide_inputs()
    wconv = torch.from_numpy(c["wconv"]).cuda()

    class Model:
        calls = 0
        def __call__(self, batch):
            half = torch.nn.functional.conv2d(raw(batch).float(), wconv, stride=2).contiguous()
            return (P._wrap(ops.upsample2x(half)),)
        def forward_half_logits(self, batch):
            Model.calls += 1
            return P._wrap(torch.nn.functional.conv2d(raw(batch).float(), wconv, stride=2).contiguous())
    rng = np.random.Generator(np.random.PCG64(91))
    imgs = [paddle.to_tensor(O.rng_normal(rng, (3, 56, 72))) for _ in range(3)]        # one even size: the fused path applies
    ori = [(56, 72)] * 3
    preds = PS.ss_inference(Model(), imgs, ori, True, None, c["stride"], c["crop"], c["nc"])
    assert Model.calls >= 1
    unfused = PS.ss_inference(lambda b: Model()(b), imgs, ori, True, None, c["stride"], c["crop"], c["nc"])
    assert len(preds) == len(unfused) == 3
    for a, b in zip(preds, unfused):
        assert a.shape == b.shape and (raw(a) == raw(b)).float().mean().item() >= 0.999


def _fake_reference_encoder_decoder(P, params):
    """A Layer with exactly the reference EncoderDecoder's parameter tree (the reference class itself is not on the GPU
    box; tests/test_reference_pin.py::test_fast_encoder_decoder_subclasses_the_reference_class covers the real one on CPU).
    Its own forward raises: the test proves the native path ran."""

    class FakeRef(P.Layer):
        def __init__(self):
            super().__init__()
            self.nhead = 8
            for key, value in params.items():
                node = self
                *path, leaf = key.split(".")
                for part in path:
                    if part not in node._modules:
                        node.add_module(part, P.Layer())
                    node = node._modules[part]
                t = torch.as_tensor(np.asarray(value)).cuda().as_subclass(P.Tensor)
                node.register_parameter(leaf, torch.nn.Parameter(t, requires_grad=True))

        def forward(self, src_feats, src_psp, src_mask=None):
            raise AssertionError("the reference forward must not run in eval mode under no_grad")

    return FakeRef


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_binding_fast_encoder_decoder_native_path(binding, dtype):
    PS, paddle = binding
    from oracle import paddle_on_torch as P
    g = load("ref_encdec_full")
    ne, nd = int(g["num_enc"]), int(g["num_dec"])
    c = G.encdec_inputs(int(g["tile"]), int(g["B"]), int(g["seed"]), ne, nd)
    Fast = PS.make_fast_encoder_decoder(_fake_reference_encoder_decoder(P, c["params"]))
    m = Fast()
    assert sorted(k for k, _ in m.named_parameters()) == list(g["keys"])
    m.eval()
    T = paddle.to_tensor
    with paddle.no_grad():
        hs, mem = m([T(f).astype(dtype) for f in c["feats"]], T(c["psp"]).astype(dtype))
    assert hs.shape == list(g["hs"].shape) and mem.shape == list(g["memory"].shape)
    tol = 5e-4 if dtype == "float32" else None
    if tol:
        assert rel_err(mem, g["memory"]) < tol and rel_err(hs, g["hs"]) < tol
    else:
        _, (rhs, rmem) = oracle_encdec_pair(c["params"], c["feats"], c["psp"], ne, nd)
        assert_bf16_parity(raw(mem).float(), g["memory"], rmem, "binding memory vs reference")
        assert_bf16_parity(raw(hs).float(), g["hs"], rhs, "binding hs vs reference")
    # parameters are aliased, not copied: an in-place update of a Paddle parameter reaches the native module
    native = m._native_module()
    key = "decoder.layers.0.norm3.bias"
    before = dict(native.named_parameters())[key].detach().clone()
    dict(m.named_parameters())[key].as_subclass(torch.Tensor).data.add_(1.0)
    assert torch.equal(dict(native.named_parameters())[key].detach(), before + 1.0)
