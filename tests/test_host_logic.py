"""Host-side logic on CPU: window planning vs the oracle's restatement of infer.py, sharding, and the N>1
partition under a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch

import oracle as O
from emrt_b200 import infer, sharding


@pytest.mark.parametrize("size,crop,stride", [(1024, 512, 384), (6000, 512, 384), (256, 512, 384), (513, 512, 384),
                                               (50, 24, 16), (70, 32, 20), (41, 24, 16), (33, 32, 20)])
def test_window_origins_equal_oracle(size, crop, stride):
    assert infer.window_origins(size, crop, stride) == O.window_origins(size, crop, stride)


def test_plan_windows_reproduces_reference_accumulation():
    """Accumulate an index-valued 'logit' with our plan and with the oracle's slide_inference: identical canvases."""
    rng = np.random.Generator(np.random.PCG64(0))
    imgs = [torch.from_numpy(O.rng_normal(rng, (3, 50, 70))), torch.from_numpy(O.rng_normal(rng, (3, 41, 33)))]
    crop, stride = (32, 24), (20, 16)
    want = O.slide_inference(lambda b: (b,), imgs, crop, stride, 3)
    plan, mh, mw = infer.plan_windows([(50, 70), (41, 33)], crop, stride)
    canvas = torch.zeros(2, 3, mh, mw)
    count = torch.zeros(2, 1, mh, mw)
    for (i, y0, x0, wh, ww) in plan:
        canvas[i, :, y0:y0 + wh, x0:x0 + ww] += imgs[i][:, y0:y0 + wh, x0:x0 + ww]
        count[i, :, y0:y0 + wh, x0:x0 + ww] += 1
    for i, (h, w) in enumerate([(50, 70), (41, 33)]):
        assert torch.equal(canvas[i:i + 1, :, :h, :w] / count[i:i + 1, :, :h, :w], want[i])


def test_plan_counts_for_baseline_configs():
    plan, _, _ = infer.plan_windows([(1024, 1024)], (512, 512), (384, 384))
    assert len(plan) == 9 and sorted({p[1] for p in plan}) == [0, 384, 512]
    plan, _, _ = infer.plan_windows([(6000, 6000)], (512, 512), (384, 384))
    assert len(plan) == 256 and max(p[1] for p in plan) == 5488


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 9, 72, 256):
        for world in (1, 2, 4, 8):
            got = []
            for r in range(world):
                b, e = sharding.shard_range(n, r, world)
                got += list(range(b, e))
            assert got == list(range(n))
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_shard_scene_rows_cover_all_label_rows_once():
    rows = O.window_origins(6000, 512, 384)
    for world in (1, 2, 4, 8):
        covered = np.zeros(6000, int)
        for r in range(world):
            own, y0, y1, halo = sharding.shard_scene_rows(rows, 512, r, world)
            covered[y0:y1] += 1
            have = set(own) | set(halo)
            for y in (y0, (y0 + y1) // 2, y1 - 1):     # every window row covering these label rows is local
                need = {k for k, o in enumerate(rows) if o <= y < o + 512}
                assert need <= have
        assert (covered == 1).all()


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_images(9, rank, world)
    t = torch.zeros(9)
    t[mine] = 1
    dist.all_reduce(t)                         # test-only check that the shards tile the image set exactly once
    areas = torch.tensor([float(len(mine)), 1.0])
    dist.all_reduce(areas)                     # the val.py:168-170 metric reduction, 3 x [num_classes] there
    q.put((rank, t.tolist(), areas.tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(60) for p in ps]
    for _, cover, areas in res:
        assert cover == [1.0] * 9 and areas == [9.0, 2.0]


@pytest.mark.parametrize("shapes", [[(64, 64), (32, 32), (16, 16)], [(8, 6), (4, 3), (2, 2)], [(47, 33), (23, 17)]])
def test_reference_points_host_mirror_equals_oracle(shapes):
    """get_reference_points (t_e_d.py:213-228) host mirror == the oracle restatement, bit for bit; a non-trivial
    valid_ratios follows the same arithmetic."""
    import numpy as np
    import oracle as O
    from emrt_b200.refpoints import reference_points_host
    got = reference_points_host(shapes)
    assert got.shape == (1, sum(h * w for h, w in shapes), len(shapes), 2)
    assert np.array_equal(got, O.encoder_reference_points(shapes, 1).numpy())
    vr = np.full((2, len(shapes), 2), 0.75, np.float32)
    got_vr = reference_points_host(shapes, vr)
    assert got_vr.shape[0] == 2 and np.allclose(got_vr, got, atol=1e-6)     # (x / (vr * W)) * vr == x / W


def _grad_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from emrt_b200.train import GradientBuckets
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(256, 288)), torch.nn.Parameter(torch.zeros(288)),
              torch.nn.Parameter(torch.zeros(256, 256)), torch.nn.Parameter(torch.zeros(7))]
    gb = GradientBuckets(params, bucket_mb=0.3)            # forces several buckets
    assert len(gb.buckets) >= 2 and gb.nbytes == 4 * sum(p.numel() for p in params)
    loss = sum(((rank + 1) * (i + 1)) * p.sum() for i, p in enumerate(params[:3]))     # params[3] never gets a grad
    loss.backward()
    assert all(p.grad.data_ptr() >= gb.buckets[0].data_ptr() for p in params[:1])      # grads live in the buckets
    gb.all_reduce()
    mean = sum(r + 1 for r in range(world)) / world
    ok = all(torch.allclose(p.grad, torch.full_like(p, mean * (i + 1))) for i, p in enumerate(params[:3]))
    ok = ok and bool((params[3].grad == 0).all())
    gb.zero()
    ok = ok and all(float(p.grad.abs().sum()) == 0 for p in params)
    out[rank] = ok
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_buckets_average_in_place():
    """cfg 4's exchange step (DataParallel grad sync, train.py:116-123,153): bucketed in-place average over 2 ranks."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_grad_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_shapes_to_host_never_returns_a_stale_entry():
    """Regression: the shape cache used to be keyed on data_ptr, so a NEW tensor that reused a freed tensor's address got
    the old shapes.  CPU tensors are read every time; equal-looking temporaries must give their own values."""
    import torch
    from emrt_b200.msda import shapes_to_host
    seen = []
    for shapes in ([(32, 32), (16, 16), (8, 8)], [(64, 64), (32, 32), (16, 16)], [(8, 6), (4, 3), (2, 2)]):
        t = torch.tensor(shapes)
        seen.append(shapes_to_host(t))
        del t
    assert seen == [((32, 32), (16, 16), (8, 8)), ((64, 64), (32, 32), (16, 16)), ((8, 6), (4, 3), (2, 2))]
    assert shapes_to_host([(5, 9)]) == ((5, 9),)


@pytest.mark.parametrize("h,w", [(64, 64), (8, 6), (5, 9)])
def test_position_embedding_host_mirror_equals_oracle(h, w):
    """PositionEmbedding (position_encoding.py:51-75, all-ones mask) host constant == the oracle restatement."""
    import numpy as np
    import oracle as O
    from emrt_b200.decoder import position_embedding_sine_host
    got = position_embedding_sine_host(h, w, 128)
    want = O.position_embedding_sine(h, w, 128).numpy().reshape(h * w, 256)
    assert got.shape == (h * w, 256) and np.abs(got - want).max() < 1e-6


def test_encoder_decoder_state_dict_keys_match_the_reference_layout():
    """The mirror's parameter names / shapes are the reference's (the oracle's generator lists them with Paddle key
    names): a checkpoint's EncoderDecoder sub-dict loads without renaming."""
    import oracle as O
    import emrt_b200
    m = emrt_b200.EncoderDecoder(110, "sine", False, (512, 1024, 2048), 3, 6, 6, 6, 256, 8, 4, 2, 1024)
    p = O.make_encoder_decoder_params(1, num_enc=4, num_dec=2)
    sd = m.state_dict()
    assert set(sd) == set(p)
    assert all(tuple(sd[k].shape) == tuple(p[k].shape) for k in sd)


def test_window_center_hint_of_reference_init():
    """The reference initialises sampling_offsets.bias to one direction per head, 1..P pixels out (t_e_d.py:47-55): the
    hint is the rounded mid-range 3.5 * direction of every (head, level)."""
    import math
    import torch
    from emrt_b200.msda import window_center_hint
    from oracle import emrt_oracle as O
    M, L_, P = 8, 3, 6
    bias = torch.from_numpy(O.msda_reset_parameters(256, M, L_, P).reshape(-1))
    hint = torch.tensor(window_center_hint(bias, M, L_, P)).view(M, L_, 2)
    for m in range(M):
        th = 2.0 * math.pi * m / M
        dx, dy = math.cos(th), math.sin(th)
        s = max(abs(dx), abs(dy))
        want = torch.tensor([dx / s * 3.5, dy / s * 3.5])
        for l in range(L_):      # (the mid-range is exactly +-3.5 on the dominant axis: either neighbour is a correct rounding)
            assert (hint[m, l].float() - want).abs().max() <= 0.5 + 1e-4, (m, l, hint[m, l], want)
    # a zero bias (no preferred direction) gives no shift; large biases are clamped
    assert window_center_hint(torch.zeros(M * L_ * P * 2), M, L_, P) == [0] * (M * L_ * 2)
    assert max(window_center_hint(torch.full((M * L_ * P * 2,), 1e4), M, L_, P)) == 100
