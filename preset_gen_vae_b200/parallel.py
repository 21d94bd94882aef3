"""Host-side data-parallel plumbing (one process per GPU, torch.distributed; NCCL on the B200 box, gloo in CPU tests).

The reference trains with single-process nn.DataParallel (train.py:95-97): every step it scatters the batch, re-broadcasts
all 241.5 MB of parameters, gathers the outputs and reduces the gradients onto the main device.  Here every rank keeps
its parameters resident, processes its own shard of the global batch with per-rank BatchNorm statistics (the semantics
of DataParallel replicas) and the only exchange is ONE all-reduce (sum, fp32) of a flat gradient buffer, pre-scaled by
1/world so that equal shards reproduce the full-batch mean gradient.

Caveat: SynthParamsLoss normalises each categorical group by the number of useful rows of the batch it
sees (loss.py:172).  With equal shards this equals the global normalisation only when every shard has the same count;
`global_useful_counts` all-reduces the 54 counts so a caller can rescale, and the difference is otherwise documented.
"""
from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int):
    """Contiguous, equal shards (the training DataLoader drops the last incomplete batch, data/build.py:64-67)."""
    if global_batch % world != 0:
        raise ValueError("global batch %d is not divisible by the world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


class FlatLayout:
    """Offsets of many tensors inside one flat fp32 buffer; every slot starts on a 16-byte boundary."""

    def __init__(self, sizes: Sequence[int], align_elems: int = 4):
        self.sizes = [int(s) for s in sizes]
        padded = [(s + align_elems - 1) // align_elems * align_elems for s in self.sizes]
        self.offsets = np.concatenate([[0], np.cumsum(padded)])[:-1].astype(np.int64)
        self.total = int(sum(padded))

    def views(self, flat: torch.Tensor, shapes: Sequence[torch.Size]) -> List[torch.Tensor]:
        return [flat[o:o + n].view(shape) for o, n, shape in zip(self.offsets, self.sizes, shapes)]

    def pack_(self, flat: torch.Tensor, tensors: Sequence[torch.Tensor], scale: float = 1.0):
        """Reference (torch) implementation of pgv_multi_pack, for CPU tests and non-CUDA tensors."""
        for o, n, t in zip(self.offsets, self.sizes, tensors):
            flat[o:o + n].copy_(t.reshape(-1))
        if scale != 1.0:
            flat.mul_(scale)
        return flat


def complement_segments(total: int, excluded: Sequence[Sequence[int]]):
    """[lo, hi) ranges of a flat buffer of `total` elements that are NOT covered by the `excluded` (offset, size) slots.
    TrainStep all-reduces the excluded slots (the FC weight gradients) early and these ranges after the step."""
    out, pos = [], 0
    for lo, n in sorted((int(o), int(n)) for o, n in excluded):
        if lo < pos or lo + n > total:
            raise ValueError("excluded slots overlap or leave the buffer")
        if lo > pos:
            out.append((pos, lo))
        pos = lo + n
    if pos < total:
        out.append((pos, total))
    return out


def merged_slot_ranges(offsets: Sequence[int], total: int, chosen: Sequence[int]):
    """[lo, hi) ranges covering the slots `chosen` (indices into `offsets`) of a flat buffer; a slot extends to the next slot's
    offset (its alignment padding belongs to it), so that neighbouring slots merge into one range - one collective per range."""
    bounds = [int(o) for o in offsets] + [int(total)]
    out = []
    for i in sorted(set(int(c) for c in chosen)):
        lo, hi = bounds[i], bounds[i + 1]
        if out and out[-1][1] == lo:
            out[-1] = (out[-1][0], hi)
        else:
            out.append((lo, hi))
    return out


def shard_of_segment(lo: int, hi: int, rank: int, world: int):
    """This rank's equal share [my_lo, my_hi) of the flat range [lo, hi); the length must divide by `world` (FlatLayout pads every
    slot to 4 x world elements for that) and shards stay 16-byte aligned."""
    n = hi - lo
    if n % world:
        raise ValueError("segment of %d elements does not split into %d equal shards" % (n, world))
    c = n // world
    return lo + rank * c, lo + (rank + 1) * c


def _native_shard_collectives(group) -> bool:
    return dist.get_backend(group) == 'nccl'


def reduce_scatter_segments_(flat: torch.Tensor, segments: Sequence[Sequence[int]], group=None):
    """In place, per segment [lo, hi): afterwards this rank's shard of the segment holds the SUM over ranks of that shard (the rest of
    the segment is unspecified).  NCCL: reduce_scatter_tensor with the output aliasing its slice of the input; other backends (gloo in
    the CPU tests) emulate it with an all-reduce."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    for lo, hi in segments:
        mlo, mhi = shard_of_segment(lo, hi, rank, world)
        if _native_shard_collectives(group):
            dist.reduce_scatter_tensor(flat[mlo:mhi], flat[lo:hi], op=dist.ReduceOp.SUM, group=group)
        else:
            total = flat[lo:hi].clone()
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
            flat[mlo:mhi].copy_(total[mlo - lo:mhi - lo])
    return flat


def all_gather_segments_(flat: torch.Tensor, segments: Sequence[Sequence[int]], group=None):
    """In place, per segment: every rank's shard is copied to all ranks (the counterpart of reduce_scatter_segments_)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    for lo, hi in segments:
        mlo, mhi = shard_of_segment(lo, hi, rank, world)
        if _native_shard_collectives(group):
            dist.all_gather_into_tensor(flat[lo:hi], flat[mlo:mhi], group=group)
        else:
            parts = [torch.empty_like(flat[mlo:mhi]) for _ in range(world)]
            dist.all_gather(parts, flat[mlo:mhi].clone(), group=group)
            flat[lo:hi].copy_(torch.cat(parts))
    return flat


def allreduce_segments_(flat: torch.Tensor, slots: Sequence[Sequence[int]], group=None, early_stream=None, early_event=None):
    """Sum over ranks of `flat`, issued as one collective per excluded slot followed by one per complement range (same order on
    every rank).  With `early_stream` / `early_event` (CUDA) the slot collectives are enqueued from that stream once the event
    has fired, i.e. possibly while the rest of the buffer is still being produced.  Returns when the caller's stream may read
    the reduced buffer."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return flat
    works = []
    if early_stream is not None:
        early_stream.wait_event(early_event)
        with torch.cuda.stream(early_stream):
            for lo, n in sorted((int(o), int(n)) for o, n in slots):
                works.append(dist.all_reduce(flat[lo:lo + n], op=dist.ReduceOp.SUM, group=group, async_op=True))
    else:
        for lo, n in sorted((int(o), int(n)) for o, n in slots):
            works.append(dist.all_reduce(flat[lo:lo + n], op=dist.ReduceOp.SUM, group=group, async_op=True))
    for lo, hi in complement_segments(flat.numel(), slots):
        works.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in works:
        w.wait()
    return flat


def allreduce_mean_(flat_grads_prescaled: torch.Tensor, group=None):
    """Sum over ranks of the flat gradient buffer.  The mean needs a 1/world factor, applied either beforehand by the caller
    (pre-scaled buffer) or afterwards (TrainStep: the fused Adam kernel's grad_scale)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grads_prescaled, op=dist.ReduceOp.SUM, group=group)
    return flat_grads_prescaled


def global_useful_counts(local_counts: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce of the per-categorical-group useful-row counts (they depend on the targets only)."""
    out = local_counts.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def max_over_ranks(value: float, device, group=None) -> float:
    """Timing helper: the slowest rank defines the step time."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
