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

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced shards; the first ``global_batch % world_size`` ranks get one extra."""
    if global_batch < 0 or world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad shard request")
    base, rem = divmod(global_batch, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def balanced_counts(global_batch: int, rates: Sequence[float]) -> List[int]:
    """Shard sizes proportional to each rank's measured rate (images/s of its own forward), summing to
    ``global_batch``.  The GPUs of one box do not run at the same speed under the 1000 W power cap (a few percent
    apart), and every step ends in a collective: with equal shards all ranks run at the pace of the slowest.  Largest
    remainders get the leftover images; ties go to the lower rank, so every rank computes the same list."""
    if global_batch < 0 or not rates or any(not (r > 0) for r in rates):
        raise ValueError("balanced_counts needs positive rates")
    total = float(sum(rates))
    exact = [global_batch * r / total for r in rates]
    counts = [int(e) for e in exact]
    order = sorted(range(len(rates)), key=lambda i: (-(exact[i] - counts[i]), i))
    for i in order[: global_batch - sum(counts)]:
        counts[i] += 1
    return counts


def rebalance(counts: Sequence[int], step_ms: float, waits_ms: Sequence[float], cap: Optional[int] = None) -> List[int]:
    """New shard sizes from one lockstep measurement: every step ends in the logits all-gather, so all ranks take
    ``step_ms`` per step and rank r spends ``waits_ms[r]`` of it inside the gather waiting for the slowest rank.  Its own
    forward therefore takes ``step_ms - waits_ms[r]`` for ``counts[r]`` images; the global batch is re-divided in
    proportion to those rates (``balanced_counts``).  Measured in the running loop itself, i.e. in the thermal / power
    state that matters (a separate calibration burst is not: profiles/r02_scaling.md).  ``cap`` bounds a shard."""
    if len(counts) != len(waits_ms) or step_ms <= 0:
        raise ValueError("rebalance needs one wait per rank and a positive step time")
    rates = [max(1, c) / max(step_ms - w, 0.25 * step_ms) for c, w in zip(counts, waits_ms)]
    total = int(sum(counts))
    new = balanced_counts(total, rates)
    if cap is not None:
        if cap * len(counts) < total:
            raise ValueError("cap too small for the global batch")
        capped = set()
        while any(c > cap for i, c in enumerate(new) if i not in capped):   # pin the ranks at the cap, re-divide the rest
            capped |= {i for i, c in enumerate(new) if c >= cap}
            free = [i for i in range(len(counts)) if i not in capped]
            rest = balanced_counts(total - cap * len(capped), [rates[i] for i in free]) if free else []
            new = [cap] * len(counts)
            for i, c in zip(free, rest):
                new[i] = c
    return new


def counts_range(counts: Sequence[int], rank: int) -> Tuple[int, int]:
    """Contiguous shard ``[start, stop)`` of ``rank`` for explicit shard sizes."""
    start = int(sum(counts[:rank]))
    return start, start + int(counts[rank])


_compact_index = {}


def _compaction_index(counts: Tuple[int, ...], per: int, device) -> torch.Tensor:
    """Row r of the compact [sum(counts), C] logits <- row of the padded [world * per, C] gather buffer."""
    key = (counts, per, str(device))
    idx = _compact_index.get(key)
    if idx is None:
        idx = torch.cat([torch.arange(n, dtype=torch.int64) + r * per for r, n in enumerate(counts)]).to(device)
        _compact_index[key] = idx
    return idx


def sharded_logits(forward_local: Callable[[torch.Tensor, torch.Tensor], None],
                   images_local: torch.Tensor, global_batch: int, num_classes: int,
                   group: Optional[dist.ProcessGroup] = None, counts: Optional[Sequence[int]] = None) -> torch.Tensor:
    """Run ``forward_local(images_local, out_slot)`` and all-gather to ``[global_batch, C]``.

    ``forward_local`` must write fp32 logits for its shard into ``out_slot`` (a view of
    the gather buffer).  Equal shards use the in-place ``all_gather_into_tensor``;
    ragged shards (``counts`` given -- see ``balanced_counts`` -- or a batch the world size does not divide) gather
    padded slots and compact them with one index_select."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if counts is not None:
        if len(counts) != world or sum(counts) != global_batch:
            raise ValueError("counts must hold one shard size per rank and sum to global_batch")
        start, stop = counts_range(counts, rank)
    else:
        start, stop = shard_range(global_batch, world, rank)
    if images_local.shape[0] != stop - start:
        raise ValueError(f"rank {rank} expected {stop - start} images, got {images_local.shape[0]}")
    dev = images_local.device
    if world == 1:
        out = torch.empty((global_batch, num_classes), dtype=torch.float32, device=dev)
        forward_local(images_local, out)
        return out
    if counts is not None and len(set(counts)) > 1:
        per = max(counts)
        padded = torch.empty((world, per, num_classes), dtype=torch.float32, device=dev)
        if stop > start:
            forward_local(images_local, padded[rank, : stop - start])
        send = padded[rank] if dist.get_backend(group) == "nccl" else padded[rank].clone()
        dist.all_gather_into_tensor(padded.view(world * per, num_classes), send, group=group)
        return padded.view(world * per, num_classes).index_select(0, _compaction_index(tuple(int(c) for c in counts), per, dev))
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
                         group: Optional[dist.ProcessGroup] = None, counts: Optional[Sequence[int]] = None):
    """Serving loop of the sharded forward, end to end: iterate this rank's HOST shards (pinned float32 tensors
    ``[B/G, *image_shape]``) and yield the gathered HOST logits ``[global_batch, C]`` of every step, in order.

    Two steps are in flight: the host->device copy of shard k+1 (copy stream) overlaps the forward + logits
    all-gather of step k (current stream), whose device->host read-back (output stream) overlaps step k+1.
    Every step's images cross PCIe, every step's gathered logits come back to the host, and the all-gather is
    inside the loop -- this is the path ``bench.py`` times as `e2e` at N > 1.  Yielded tensors are pinned staging
    buffers re-used two steps later."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    uneven = counts is not None and len(set(counts)) > 1
    if counts is not None:
        if len(counts) != world or sum(counts) != global_batch:
            raise ValueError("counts must hold one shard size per rank and sum to global_batch")
        start, stop = counts_range(counts, rank)
    else:
        start, stop = shard_range(global_batch, world, rank)
        if global_batch % world != 0:
            raise ValueError("sharded_apply_stream needs equal shards or explicit counts")
    local = stop - start
    per = max(counts) if uneven else local
    index = _compaction_index(tuple(int(c) for c in counts), per, device) if uneven else None
    cuda = device.type == "cuda"
    main = torch.cuda.current_stream(device) if cuda else None
    copy_s = torch.cuda.Stream(device) if cuda else None
    out_s = torch.cuda.Stream(device) if cuda else None
    img = [torch.empty((local,) + tuple(image_shape), dtype=torch.float32, device=device) for _ in range(2)]
    gath = [torch.empty((world * per if uneven else global_batch, num_classes), dtype=torch.float32, device=device)
            for _ in range(2)]
    compact = [torch.empty((global_batch, num_classes), dtype=torch.float32, device=device) for _ in range(2)] if uneven else gath
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
        if uneven:      # padded slots of `per` rows, compacted on the device before the read-back
            forward_local(img[s], gath[s][rank * per: rank * per + local])
            send = gath[s][rank * per: (rank + 1) * per]
            dist.all_gather_into_tensor(gath[s], send if nccl else send.clone(), group=group)
            torch.index_select(gath[s], 0, index, out=compact[s])
        else:
            slot = gath[s][start:stop]
            forward_local(img[s], slot)
            if world > 1:
                dist.all_gather_into_tensor(gath[s], slot if nccl else slot.clone(), group=group)
        if cuda:
            fwd_done[s].record(main)
            with torch.cuda.stream(out_s):
                out_s.wait_event(fwd_done[s])
                host[s].copy_(compact[s], non_blocking=True)
                d2h_done[s].record(out_s)
        else:
            host[s].copy_(compact[s])
        used[s] = True
        pending.append(s)
        k += 1
    while pending:
        t = pending.pop(0)
        if cuda:
            d2h_done[t].synchronize()
        yield host[t]
