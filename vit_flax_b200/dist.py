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


def sharded_apply_stream(forward_local: Callable[[torch.Tensor, torch.Tensor], None], host_batches,
                         global_batch: int, num_classes: int, image_shape: Tuple[int, ...], device: torch.device,
                         group: Optional[dist.ProcessGroup] = None):
    """Serving loop of the sharded forward, end to end: iterate this rank's HOST shards (pinned float32 tensors
    ``[B/G, *image_shape]``) and yield the gathered HOST logits ``[global_batch, C]`` of every step, in order.

    Two steps are in flight: the host->device copy of shard k+1 (copy stream) overlaps the forward + logits
    all-gather of step k (current stream), whose device->host read-back (output stream) overlaps step k+1.
    Every step's images cross PCIe, every step's gathered logits come back to the host, and the all-gather is
    inside the loop -- this is the path ``bench.py`` times as `e2e` at N > 1.  Yielded tensors are pinned staging
    buffers re-used two steps later."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    start, stop = shard_range(global_batch, world, rank)
    if global_batch % world != 0:
        raise ValueError("sharded_apply_stream needs equal shards (global_batch divisible by the world size)")
    local = stop - start
    cuda = device.type == "cuda"
    main = torch.cuda.current_stream(device) if cuda else None
    copy_s = torch.cuda.Stream(device) if cuda else None
    out_s = torch.cuda.Stream(device) if cuda else None
    img = [torch.empty((local,) + tuple(image_shape), dtype=torch.float32, device=device) for _ in range(2)]
    gath = [torch.empty((global_batch, num_classes), dtype=torch.float32, device=device) for _ in range(2)]
    host = [torch.empty((global_batch, num_classes), dtype=torch.float32, pin_memory=cuda) for _ in range(2)]
    ev = lambda: torch.cuda.Event() if cuda else None
    h2d_done, fwd_done, d2h_done = [ev(), ev()], [ev(), ev()], [ev(), ev()]
    used = [False, False]
    pending = []
    nccl = dist.is_initialized() and world > 1 and dist.get_backend(group) == "nccl"
    k = 0
    for shard in host_batches:
        if tuple(shard.shape) != tuple(img[0].shape):
            raise ValueError(f"rank {rank} expected a shard of shape {tuple(img[0].shape)}, got {tuple(shard.shape)}")
        s = k & 1
        if len(pending) == 2:
            t = pending.pop(0)
            if cuda:
                d2h_done[t].synchronize()
            yield host[t]
        if cuda:
            with torch.cuda.stream(copy_s):
                if used[s]:
                    copy_s.wait_event(fwd_done[s])          # the forward two steps ago has read this buffer
                img[s].copy_(shard, non_blocking=True)
                h2d_done[s].record(copy_s)
            main.wait_event(h2d_done[s])
            if used[s]:
                main.wait_event(d2h_done[s])                # its gathered logits have left for the host
        else:
            img[s].copy_(shard)
        slot = gath[s][start:stop]
        forward_local(img[s], slot)
        if world > 1:
            dist.all_gather_into_tensor(gath[s], slot if nccl else slot.clone(), group=group)
        if cuda:
            fwd_done[s].record(main)
            with torch.cuda.stream(out_s):
                out_s.wait_event(fwd_done[s])
                host[s].copy_(gath[s], non_blocking=True)
                d2h_done[s].record(out_s)
        else:
            host[s].copy_(gath[s])
        used[s] = True
        pending.append(s)
        k += 1
    while pending:
        t = pending.pop(0)
        if cuda:
            d2h_done[t].synchronize()
        yield host[t]
