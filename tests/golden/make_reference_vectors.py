"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCES (imported in place from
/root/reference by oracle/run_reference.py) on oracle/paddle_on_torch.py, the torch-CPU mapping of the Paddle
operators they call.  These are outputs of the reference's code for seeded inputs and weights: the functions of
SURVEY.md §8(a) rows a1-a8 plus the §8(f) rows that are built (encoder layer, decoder, EncoderDecoder).

What this pins: the reference's layouts, reshapes / transposes / splits, level and window loops, parameter shapes and
state-dict keys, initialisation, and the composition of operators.  What it cannot pin: PaddlePaddle's own kernels
(the operator semantics assumed are listed in oracle/paddle_on_torch.py); a real Paddle run remains the final check
(INTEGRATION.md).

Run here (the build container; needs /root/reference):   python tests/golden/make_reference_vectors.py
Inputs and weights are regenerated in the tests from the seeds below (numpy PCG64 via oracle.make_*_params /
rng_normal); each file stores a float64 checksum of them so that generator drift is detected, not silently compared.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle.emrt_oracle as O  # noqa: E402
from oracle import run_reference as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
f32 = lambda t: np.ascontiguousarray(torch.as_tensor(t).detach().float().numpy())


def checksum(arrays):
    return np.float64(sum(float(np.asarray(a, dtype=np.float64).sum()) + 1e-3 * float(np.abs(np.asarray(a, dtype=np.float64)).sum())
                          for a in arrays))


# seeded inputs shared with the tests (tests/refcases.py imports these builders) -------------------------------------
def msda_inputs():
    rng = np.random.Generator(np.random.PCG64(7))
    shapes = [(8, 6), (4, 3), (2, 2)]                  # non-square levels
    B, C, M, P, Lq = 2, 256, 8, 6, 37                  # Lq != Lv (decoder-style), real channel / head counts
    _, Lv = O.level_tables(shapes)
    params = O.make_msda_params(99, C, M, len(shapes), P, offset_std=0.15)
    q = O.rng_normal(rng, (B, Lq, C))
    v = O.rng_normal(rng, (B, Lv, C))
    ref = rng.uniform(0, 1, size=(B, Lq, len(shapes), 2)).astype(np.float32)
    mask = (rng.uniform(size=(B, Lv)) > 0.1).astype(np.float32)
    return dict(shapes=shapes, params=params, query=q, value=v, ref=ref, mask=mask, M=M, P=P, C=C)


def core_inputs():
    rng = np.random.Generator(np.random.PCG64(8))
    shapes = [(16, 16), (8, 8), (4, 4)]
    B, M, D, P, Lq = 2, 8, 32, 6, 50
    _, Lv = O.level_tables(shapes)
    value = O.rng_normal(rng, (B, Lv, M, D))
    loc = rng.uniform(-0.25, 1.25, size=(B, Lq, M, len(shapes), P, 2)).astype(np.float32)   # ~30 % of samples leave the map
    attn = rng.uniform(0, 1, size=(B, Lq, M, len(shapes), P)).astype(np.float32)
    attn /= attn.reshape(B, Lq, M, -1).sum(-1)[..., None, None]
    return dict(shapes=shapes, value=value, loc=loc, attn=attn)


def encdec_inputs(tile, B, seed, num_enc, num_dec):
    params = O.make_encoder_decoder_params(seed, num_enc=num_enc, num_dec=num_dec)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    feats = [O.rng_normal(rng, (B, c, tile // s, tile // s), 0.5) for c, s in zip((512, 1024, 2048), (8, 16, 32))]
    psp = O.rng_normal(rng, (B, 256, 110), 0.5)
    return dict(params=params, feats=feats, psp=psp)


def masked_src_mask():
    """[2, 64, 64] float32, 1 = valid: image 0 is padded on the right / bottom (valid 48 rows x 40 columns), image 1 is full."""
    m = np.zeros((2, 64, 64), np.float32)
    m[0, :48, :40] = 1.0
    m[1] = 1.0
    return m


def mha_inputs():
    rng = np.random.Generator(np.random.PCG64(21))
    params = {k[len("decoder.layers.0.self_attn."):]: v for k, v in O.make_encoder_decoder_params(20, num_enc=0, num_dec=1).items()
              if k.startswith("decoder.layers.0.self_attn.")}
    tgt = O.rng_normal(rng, (2, 110, 256))
    pos = O.rng_normal(rng, (2, 110, 256), 0.5)
    return dict(params=params, tgt=tgt, pos=pos)


def uphead_inputs():
    rng = np.random.Generator(np.random.PCG64(31))
    nc = 6
    p = {}
    for i, (co, ci, k) in enumerate(((256, 256, 3), (256, 256, 3), (256, 256, 3), (nc, 256, 1))):
        p[f"conv_{i}.weight"] = O.rng_uniform(rng, (co, ci, k, k), (6.0 / (ci * k * k + co * k * k)) ** 0.5)
        p[f"conv_{i}.bias"] = O.rng_uniform(rng, (co,), 0.1)
    for i in range(3):
        p[f"syncbn_fc_{i}.weight"] = rng.uniform(0.5, 1.5, size=(256,)).astype(np.float32)
        p[f"syncbn_fc_{i}.bias"] = O.rng_normal(rng, (256,), 0.1)
        p[f"syncbn_fc_{i}._mean"] = O.rng_normal(rng, (256,), 0.1)
        p[f"syncbn_fc_{i}._variance"] = rng.uniform(0.5, 1.5, size=(256,)).astype(np.float32)
    x = O.rng_normal(rng, (2, 256, 6, 5))
    return dict(params=p, x=x, nc=nc)


def slide_inputs():
    rng = np.random.Generator(np.random.PCG64(11))
    nc, crop, stride = 6, (32, 24), (20, 16)            # (w, h) order like the reference
    imgs = [O.rng_normal(rng, (3, 50, 70)), O.rng_normal(rng, (3, 42, 34))]
    wconv = O.rng_normal(rng, (nc, 3, 2, 2), 0.5)       # toy model: stride-2 2x2 conv -> half-res logits -> x2 bilinear
    ori = [(50, 70), (60, 51)]
    return dict(nc=nc, crop=crop, stride=stride, imgs=imgs, wconv=wconv, ori=ori)


def area_inputs():
    rng = np.random.Generator(np.random.PCG64(41))
    nc = 7
    pred = rng.integers(0, nc, size=(1, 1, 64, 80)).astype(np.int32)
    label = rng.integers(0, nc, size=(1, 1, 64, 80)).astype(np.int64)
    label[rng.uniform(size=label.shape) < 0.05] = 255
    return dict(nc=nc, pred=pred, label=label)


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ref = R.load()
    import paddle                                        # the shim (oracle/paddle_on_torch.py)
    T = paddle.to_tensor
    files = {}

    # a1: construction + _reset_parameters (t_e_d.py:22-63)
    torch.manual_seed(0)
    m0 = ref.ted.MSDeformableAttention(256, 8, 3, 6)
    sd0 = {k: f32(v) for k, v in m0.state_dict().items()}
    files["ref_msda_init"] = dict(**{"shape." + k: np.array(v.shape) for k, v in sd0.items()},
                                  sampling_offsets_bias=sd0["sampling_offsets.bias"],
                                  sampling_offsets_weight_absmax=np.float32(np.abs(sd0["sampling_offsets.weight"]).max()),
                                  attention_weights_absmax=np.float32(max(np.abs(sd0["attention_weights.weight"]).max(),
                                                                          np.abs(sd0["attention_weights.bias"]).max())),
                                  value_proj_bias_absmax=np.float32(np.abs(sd0["value_proj.bias"]).max()),
                                  value_proj_weight_absmax=np.float32(np.abs(sd0["value_proj.weight"]).max()),
                                  xavier_bound=np.float32((6.0 / 512) ** 0.5))

    # a2: MSDeformableAttention.forward (t_e_d.py:65-107)
    c = msda_inputs()
    m = R.load_params(ref.ted.MSDeformableAttention(c["C"], c["M"], len(c["shapes"]), c["P"]), c["params"])
    out = m(T(c["query"]), T(c["ref"]), T(c["value"]), T(c["shapes"], dtype="int64"), T(c["mask"]))
    files["ref_msda"] = dict(out=f32(out), check=checksum([c["query"], c["value"], c["ref"], c["mask"], *c["params"].values()]))

    # a3: deformable_attention_core_func (utils.py:64-97)
    c = core_inputs()
    out = ref.utils.deformable_attention_core_func(T(c["value"]), T(c["shapes"], dtype="int64"), T(c["loc"]), T(c["attn"]))
    files["ref_core"] = dict(out=f32(out), check=checksum([c["value"], c["loc"], c["attn"]]))

    # a4: TransformerEncoder.get_reference_points (t_e_d.py:213-228), valid_ratios == 1
    for name, shapes in (("sq", [(32, 32), (16, 16), (8, 8)]), ("rect", [(8, 6), (4, 3), (2, 2)])):
        rp = ref.ted.TransformerEncoder.get_reference_points(T(shapes, dtype="int64"), paddle.ones([2, 3, 2]))
        files.setdefault("ref_refpoints", {})[name] = f32(rp)
        files["ref_refpoints"][name + "_shapes"] = np.array(shapes)
    # position embedding (position_encoding.py:51-75) as EncoderDecoder builds it (t_e_d.py:404-406)
    pe = ref.position_encoding.PositionEmbedding(128, normalize=True, embed_type="sine", offset=-0.5)
    files["ref_posembed"] = dict(pos_8x6=f32(pe(paddle.ones([1, 8, 6], dtype="bool"))),
                                 pos_16x16=f32(pe(paddle.ones([1, 16, 16], dtype="bool"))))

    # f3: MultiHeadAttention (layers.py:144-311), self-attention use of t_e_d.py:283-284
    c = mha_inputs()
    mha = R.load_params(ref.layers.MultiHeadAttention(256, 8, dropout=0.1), c["params"])
    q = T(c["tgt"] + c["pos"])
    files["ref_mha"] = dict(out=f32(mha(q, q, value=T(c["tgt"]))), check=checksum([c["tgt"], c["pos"], *c["params"].values()]))

    # f1-f3: encoder layer, decoder layer, whole EncoderDecoder exactly as EMRT constructs it (paddle_EMRT.py:241-249)
    for tag, tile, B, seed, ne, nd in (("small", 64, 2, 50, 2, 1), ("full", 128, 1, 60, 4, 2)):
        c = encdec_inputs(tile, B, seed, ne, nd)
        model = ref.ted.EncoderDecoder(hidden_dim=256, dim_feedforward=1024, backbone_num_channels=[512, 1024, 2048],
                                       dropout=0.1, activation="relu", num_feature_levels=3, nhead=8,
                                       num_encoder_layers=ne, num_decoder_layers=nd, num_encoder_points=6,
                                       num_decoder_points=6, nclass=6)
        R.load_params(model, c["params"])
        hs, memory = model([T(f) for f in c["feats"]], T(c["psp"]))
        files["ref_encdec_" + tag] = dict(hs=f32(hs), memory=f32(memory), tile=np.int64(tile), B=np.int64(B), seed=np.int64(seed),
                                          num_enc=np.int64(ne), num_dec=np.int64(nd),
                                          keys=np.array(sorted(model.state_dict().keys())),
                                          check=checksum([*c["feats"], c["psp"], *[v.numpy() for v in c["params"].values()]]))

    # the masked path (t_e_d.py:408-415,440-447,466-467: src_mask -> per-level nearest mask, valid ratios, masked sine embedding,
    # scaled reference points, value masking); EMRT itself never passes a mask (paddle_EMRT.py:265)
    c = encdec_inputs(64, 2, 70, 2, 1)
    model = ref.ted.EncoderDecoder(hidden_dim=256, dim_feedforward=1024, backbone_num_channels=[512, 1024, 2048], dropout=0.1,
                                   activation="relu", num_feature_levels=3, nhead=8, num_encoder_layers=2, num_decoder_layers=1,
                                   num_encoder_points=6, num_decoder_points=6, nclass=6)
    R.load_params(model, c["params"])
    hs, memory = model([T(f) for f in c["feats"]], T(c["psp"]), T(masked_src_mask()))
    files["ref_encdec_masked"] = dict(hs=f32(hs), memory=f32(memory), src_mask=masked_src_mask(),
                                      check=checksum([*c["feats"], c["psp"], *[v.numpy() for v in c["params"].values()]]))

    # a5: UpHead (paddle_EMRT.py:115-181) as EMRT builds it (:197-198); the logits before the last x2 are captured
    c = uphead_inputs()
    up = R.load_params(ref.emrt.UpHead(embed_dim=256, num_conv=3, num_upsample_layer=1, align_corners=False, num_classes=c["nc"]),
                       c["params"])
    grabbed = {}
    h = up.conv_3.register_forward_hook(lambda mod, inp, out: grabbed.__setitem__("half", out))
    full = up(T(c["x"]))
    h.remove()
    files["ref_uphead"] = dict(half=f32(grabbed["half"]), full=f32(full), check=checksum([c["x"], *c["params"].values()]))

    # a6 + a7: slide_inference / ss_inference (src/api/infer.py:22-157) around a toy model
    c = slide_inputs()
    import paddle.nn.functional as F

    def model(batch):
        half = torch.nn.functional.conv2d(torch.as_tensor(batch).as_subclass(torch.Tensor), torch.from_numpy(c["wconv"]), stride=2)
        return (F.interpolate(T(half), scale_factor=2, mode="bilinear", align_corners=False),)
    imgs = [T(i) for i in c["imgs"]]
    logits = ref.infer.slide_inference(model, imgs, c["crop"], c["stride"], c["nc"])
    preds = ref.infer.ss_inference(model, imgs, c["ori"], True, None, c["stride"], c["crop"], c["nc"])
    files["ref_slide"] = dict(logit0=f32(logits[0]), logit1=f32(logits[1]),
                              pred0=np.asarray(torch.as_tensor(preds[0]).numpy()), pred1=np.asarray(torch.as_tensor(preds[1]).numpy()),
                              check=checksum([*c["imgs"], c["wconv"]]))
    assert files["ref_slide"]["pred0"].dtype == np.int32 and files["ref_slide"]["pred0"].shape == (1, 1, 50, 70)

    # a8: metrics.calculate_area (src/utils/metrics.py:20-69)
    c = area_inputs()
    ia, pa, la = ref.metrics.calculate_area(T(c["pred"]), T(c["label"]), c["nc"])
    files["ref_area"] = dict(intersect=f32(ia), pred=f32(pa), label=f32(la), check=checksum([c["pred"], c["label"]]))

    for name, arrays in files.items():
        path = os.path.join(OUT, name + ".npz")
        if os.path.exists(path) and "--all" not in sys.argv:      # committed vectors stay byte-identical unless --all
            continue
        np.savez_compressed(path, **arrays)
    for f in sorted(os.listdir(OUT)):
        if f.startswith("ref_"):
            print(f"{f:28s} {os.path.getsize(os.path.join(OUT, f)) / 1024:8.1f} KB")


if __name__ == "__main__":
    main()
