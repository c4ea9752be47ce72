"""Data-parallel training step on N GPUs: per-rank train_forward + backward on its own batch, then ONE
all-reduce of the contiguous gradient buffer (vit_flax_b200.dist.all_reduce_grads).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        profiles/run_train_dp.py [batch_per_gpu]"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from _util import C2  # noqa: E402
from vit_flax_b200 import init_params, perturb_params  # noqa: E402
from vit_flax_b200.dist import all_reduce_grads  # noqa: E402
from vit_flax_b200.engine import Engine  # noqa: E402

os.environ.setdefault("NCCL_DEBUG", "WARN")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = Engine(precision="fp16", max_batch=B, device=local, **C2)
eng.load_params(perturb_params(init_params(seed=1, **C2), seed=2))
dev = torch.device("cuda", local)


def data(r):
    g = torch.Generator(device=dev).manual_seed(100 + r)
    return (torch.randn((B, 224, 224, 3), device=dev, generator=g), torch.randn((B, 1000), device=dev, generator=g) / (B * world))


x, dl = data(rank)
logits = torch.empty((B, 1000), device=dev)


def step():
    eng.train_forward(x, out=logits)
    eng.backward(dl)
    return all_reduce_grads(eng.grads_flat())


g = step().clone()
# check on rank 0: the all-reduced buffer equals the sum of every rank's local gradient, recomputed here
if rank == 0:
    want = torch.zeros_like(g, dtype=torch.float64)
    for r in range(world):
        xr, dr = data(r)
        eng.train_forward(xr, out=logits)
        eng.backward(dr)
        want += eng.grads_flat().double()
    err = float((g.double() - want).abs().max() / want.abs().max())
    print(f"all-reduced gradients vs sum of per-rank gradients: max rel err {err:.2e} over {g.numel()} floats", flush=True)
    assert err < 5e-3
for _ in range(2):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
n = 6
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"ViT-B/16 training step, {world} GPU(s) x batch {B}: {t.item():.2f} ms/step = {B * world / t.item() * 1e3:.0f} images/s "
          f"(max over ranks, gradient all-reduce of {g.numel() * 4 / 1e6:.0f} MB included)", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
