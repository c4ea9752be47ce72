"""N>1 host logic on CPU: world_size-2 gloo -- shard bounds and the logits all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vit_flax_b200.dist import shard_range, sharded_apply_stream, sharded_logits


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
        if global_batch % world == 0:
            # the serving loop (host shards in, gathered host logits out, two steps in flight): 5 steps, each with
            # its own inputs, must come back in order
            steps = [full[s:e] * (i + 1) for i in range(5)]
            outs = [y.clone() for y in sharded_apply_stream(forward_local, iter(steps), global_batch, classes, (12,),
                                                            torch.device("cpu"))]
            ok = ok and len(outs) == 5 and all(torch.allclose(y, (full * (i + 1)) @ w) for i, y in enumerate(outs))
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
