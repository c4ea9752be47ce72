"""Batch-sharded forward (and data-parallel backward) across the GPUs of one node.

Images are independent (no op of vit.py mixes the batch axis), so rank r takes
the contiguous shard ``[start, stop)`` of the global batch, weights are
replicated, and the ONLY exchange step is an all-gather of the fp32 logits.
One process per GPU; ``torch.distributed`` is the plumbing (NCCL over NVLink on
GPUs, gloo in the CPU tests).  The local forward writes its logits straight
into its slot of the gather buffer, so the all-gather runs in place with no
staging copy.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced shards; the first ``global_batch % world_size`` ranks get one extra."""
    if global_batch < 0 or world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad shard request")
    base, rem = divmod(global_batch, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def sharded_logits(forward_local: Callable[[torch.Tensor, torch.Tensor], None],
                   images_local: torch.Tensor, global_batch: int, num_classes: int,
                   group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Run ``forward_local(images_local, out_slot)`` and all-gather to ``[global_batch, C]``.

    ``forward_local`` must write fp32 logits for its shard into ``out_slot`` (a view of
    the gather buffer).  Equal shards use the in-place ``all_gather_into_tensor``;
    ragged shards fall back to padded slots."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    start, stop = shard_range(global_batch, world, rank)
    if images_local.shape[0] != stop - start:
        raise ValueError(f"rank {rank} expected {stop - start} images, got {images_local.shape[0]}")
    dev = images_local.device
    if world == 1:
        out = torch.empty((global_batch, num_classes), dtype=torch.float32, device=dev)
        forward_local(images_local, out)
        return out
    if global_batch % world == 0:
        out = torch.empty((global_batch, num_classes), dtype=torch.float32, device=dev)
        slot = out[start:stop]
        forward_local(images_local, slot)
        # NCCL all-gathers in place when the send buffer is the rank's own slot; gloo wants
        # a distinct input buffer.
        send = slot if dist.get_backend(group) == "nccl" else slot.clone()
        dist.all_gather_into_tensor(out, send, group=group)
        return out
    per = -(-global_batch // world)
    padded = torch.zeros((world, per, num_classes), dtype=torch.float32, device=dev)
    if stop > start:
        forward_local(images_local, padded[rank, : stop - start])
    send = padded[rank] if dist.get_backend(group) == "nccl" else padded[rank].clone()
    dist.all_gather_into_tensor(padded.view(world * per, num_classes), send, group=group)
    rows = [padded[r, : shard_range(global_batch, world, r)[1] - shard_range(global_batch, world, r)[0]]
            for r in range(world)]
    return torch.cat(rows, dim=0)


def all_reduce_grads(grads_flat: torch.Tensor, group: Optional[dist.ProcessGroup] = None, average: bool = False) -> torch.Tensor:
    """Data-parallel backward: every rank ran ``train_forward`` / ``backward`` on its batch shard; the
    parameter gradients of the global batch are their SUM (the VJP is additive over the batch), so the
    one exchange step is an all-reduce of the contiguous gradient buffer (``Engine.grads_flat()``,
    346 MB for ViT-B/16) -- a single NCCL collective, reduced in the NVSwitch when NVLS is up.
    In place; ``average`` divides by the world size (a mean-over-global-batch loss whose cotangent
    each rank scaled by its LOCAL batch)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grads_flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            grads_flat.div_(dist.get_world_size(group))
    return grads_flat
