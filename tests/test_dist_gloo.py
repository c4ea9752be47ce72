"""N>1 host logic on CPU: world_size-2 gloo -- shard bounds and the logits all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vit_flax_b200.dist import balanced_counts, counts_range, rebalance, shard_range, sharded_apply_stream, sharded_logits


def test_shard_range_partitions_the_batch():
    for b in (0, 1, 7, 8, 256, 2048):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(b, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_balanced_counts():
    assert balanced_counts(2048, [1.0] * 8) == [256] * 8
    c = balanced_counts(512, [1.0, 0.96])                       # a GPU 4 % slower gets 4 % fewer images
    assert sum(c) == 512 and c == [261, 251]
    c = balanced_counts(2048, [100, 101, 99, 100, 97, 103, 100, 100])
    assert sum(c) == 2048 and max(c) - min(c) <= 16 and c[5] == max(c) and c[4] == min(c)
    assert balanced_counts(7, [1, 1, 1]) == [3, 2, 2]           # leftovers by largest remainder, ties to the lower rank
    assert [counts_range([3, 2, 2], r) for r in range(3)] == [(0, 3), (3, 5), (5, 7)]
    with pytest.raises(ValueError):
        balanced_counts(8, [1.0, 0.0])
    # lockstep measurement -> new shards: rank 0 idles 0.4 ms of a 10 ms step, i.e. it is 4 % faster
    c = rebalance([256, 256], 10.0, [0.42, 0.02])
    assert sum(c) == 512 and c[0] > c[1] and c == [261, 251]
    assert rebalance(c, 9.8, [0.02, 0.02]) == c                 # balanced: stays put
    assert rebalance([256] * 4, 10.0, [0.0, 3.0, 0.0, 0.0]) == [231, 331, 231, 231]
    assert rebalance([256] * 4, 10.0, [0.0, 3.0, 0.0, 0.0], cap=260) == [255, 260, 255, 254]   # capped rank pinned, rest re-divided


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, global_batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        classes = 5
        full = torch.arange(global_batch * 12, dtype=torch.float32).view(global_batch, 12)
        w = torch.linspace(-1, 1, 12 * classes).view(12, classes)
        s, e = shard_range(global_batch, world, rank)

        def forward_local(x, out):      # stand-in for Engine.forward: writes into its slot
            out.copy_(x @ w)

        got = sharded_logits(forward_local, full[s:e], global_batch, classes)
        ok = torch.allclose(got, full @ w)
        # explicit, uneven shard sizes (speed-balanced sharding): padded gather + one compaction
        counts = [global_batch - global_batch // 3, global_batch // 3]
        s2, e2 = counts_range(counts, rank)
        got2 = sharded_logits(forward_local, full[s2:e2], global_batch, classes, counts=counts)
        ok = ok and torch.allclose(got2, full @ w) and tuple(got2.shape) == (global_batch, classes)
        if global_batch % world == 0:
            # the serving loop (host shards in, gathered host logits out, two steps in flight): 5 steps, each with
            # its own inputs, must come back in order
            steps = [full[s:e] * (i + 1) for i in range(5)]
            outs = [y.clone() for y in sharded_apply_stream(forward_local, iter(steps), global_batch, classes, (12,),
                                                            torch.device("cpu"))]
            ok = ok and len(outs) == 5 and all(torch.allclose(y, (full * (i + 1)) @ w) for i, y in enumerate(outs))
            steps = [full[s2:e2] * (i + 1) for i in range(4)]      # the same loop over speed-balanced (uneven) shards
            outs = [y.clone() for y in sharded_apply_stream(forward_local, iter(steps), global_batch, classes, (12,),
                                                            torch.device("cpu"), counts=counts)]
            ok = ok and len(outs) == 4 and all(torch.allclose(y, (full * (i + 1)) @ w) for i, y in enumerate(outs))
        q.put((rank, ok, tuple(got.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("global_batch", [8, 7])
def test_sharded_logits_world2_gloo(global_batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, global_batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, (global_batch, 5)), (1, True, (global_batch, 5))]


def _grad_worker(rank, world, port, q):
    """Data-parallel backward on CPU: each rank differentiates the oracle on its batch shard; the
    all-reduced gradient buffer must equal the gradient of the whole batch (the VJP is additive)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import vit_torch
        from vit_flax_b200 import init_params, perturb_params
        from vit_flax_b200.dist import all_reduce_grads
        from vit_flax_b200.params import flatten_params
        cfg = dict(image_size=16, patch_size=8, num_classes=8, dim=64, depth=1, heads=2, mlp_dim=64)
        v = perturb_params(init_params(seed=1, **cfg), seed=2)
        rng = np.random.default_rng(3)
        img = rng.standard_normal((6, 16, 16, 3)).astype(np.float32)
        dl = rng.standard_normal((6, 8))
        s, e = shard_range(6, world, rank)
        _, g_local = vit_torch.vit_vjp(v, img[s:e], dl[s:e], **cfg)
        keys = sorted(flatten_params({"params": g_local}))
        flat = torch.cat([torch.as_tensor(flatten_params({"params": g_local})[k]).reshape(-1) for k in keys])
        all_reduce_grads(flat)
        _, g_all = vit_torch.vit_vjp(v, img, dl, **cfg)
        want = torch.cat([torch.as_tensor(flatten_params({"params": g_all})[k]).reshape(-1) for k in keys])
        q.put((rank, bool(torch.allclose(flat, want, rtol=1e-9, atol=1e-12))))
    finally:
        dist.destroy_process_group()


def test_all_reduce_grads_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
