"""CPU: the data-parallel host logic with two gloo ranks (the N > 1 path without GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from preset_gen_vae_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                                     # identical initial weights on every rank
        model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(8, 7, generator=g), torch.randn(8, 3, generator=g)
        lo, hi = parallel.shard_range(8, rank, world)
        loss = torch.nn.functional.mse_loss(model(x[lo:hi]), y[lo:hi])       # per-rank mean loss on the shard
        loss.backward()
        params = list(model.parameters())
        layout = parallel.FlatLayout([p.numel() for p in params])
        flat = torch.zeros(layout.total)
        layout.pack_(flat, [p.grad for p in params], scale=1.0 / world)
        parallel.allreduce_mean_(flat)
        # full-batch reference on every rank
        ref = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        ref.load_state_dict(model.state_dict())
        torch.nn.functional.mse_loss(ref(x), y).backward()
        err = max(float((v - p.grad).abs().max()) for v, p in zip(layout.views(flat, [p.shape for p in params]), ref.parameters()))
        # the segmented all-reduce (two "early" slots + the ranges around them) gives the same buffer as one all-reduce
        seg = torch.arange(40.0) * (rank + 1)
        parallel.allreduce_segments_(seg, [(24, 8), (4, 12)])
        seg_err = float((seg - torch.arange(40.0) * 3).abs().max())
        # reduce-scatter -> update of this rank's shard only -> all-gather == all-reduce -> full update (TrainStep._exchange_and_update)
        lay = parallel.FlatLayout([5, 9, 3, 6], align_elems=4 * world)
        segs = parallel.merged_slot_ranges(lay.offsets, lay.total, [0, 1]) + parallel.merged_slot_ranges(lay.offsets, lay.total, [3])
        g1 = torch.Generator().manual_seed(10 + rank)
        grads, params = torch.randn(lay.total, generator=g1), torch.arange(float(lay.total))
        want = grads.clone()
        dist.all_reduce(want)
        want_p = params - 0.1 * want
        parallel.reduce_scatter_segments_(grads, segs)
        for lo, hi in segs:
            mlo, mhi = parallel.shard_of_segment(lo, hi, rank, world)
            params[mlo:mhi] -= 0.1 * grads[mlo:mhi]              # "Adam" on the shard
        parallel.all_gather_segments_(params, segs)
        covered = torch.zeros(lay.total, dtype=torch.bool)
        for lo, hi in segs:
            covered[lo:hi] = True
        shard_err = float((params - want_p)[covered].abs().max()) + float((params - torch.arange(float(lay.total)))[~covered].abs().max())
        counts = parallel.global_useful_counts(torch.tensor([3.0 + rank, 4.0]))
        slow = parallel.max_over_ranks(1.0 + rank, 'cpu')
        ret[rank] = (err, counts.tolist(), slow, layout.offsets.tolist(), seg_err, shard_err)
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert len(ret) == 2
        for rank in range(world):
            err, counts, slow, offsets, seg_err, shard_err = ret[rank]
            assert seg_err == 0.0 and shard_err < 1e-6
            assert err < 1e-6                                    # equal shards: mean of shard means == full-batch mean
            assert counts == [7.0, 8.0] and slow == 2.0
            assert all(o % 4 == 0 for o in offsets)              # 16-byte aligned slots


def test_merged_slot_ranges_cover_padding_and_merge_neighbours():
    layout = parallel.FlatLayout([5, 8, 3, 4])                 # offsets 0, 8, 16, 20; total 24
    total = layout.total
    assert parallel.merged_slot_ranges(layout.offsets, total, [0, 1, 3]) == [(0, 16), (20, 24)]
    assert parallel.merged_slot_ranges(layout.offsets, total, [2]) == [(16, 20)]
    early, late = [1, 2], [0, 3]
    cover = sorted(parallel.merged_slot_ranges(layout.offsets, total, early) + parallel.merged_slot_ranges(layout.offsets, total, late))
    assert cover[0][0] == 0 and cover[-1][1] == total and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))


def test_complement_segments():
    assert parallel.complement_segments(40, [(24, 8), (4, 12)]) == [(0, 4), (16, 24), (32, 40)]
    assert parallel.complement_segments(10, [(0, 10)]) == [] and parallel.complement_segments(10, []) == [(0, 10)]
    assert parallel.complement_segments(10, [(0, 3), (3, 7)]) == []
    with pytest.raises(ValueError):
        parallel.complement_segments(10, [(0, 6), (4, 2)])
    with pytest.raises(ValueError):
        parallel.complement_segments(10, [(8, 4)])


def test_shard_range_and_layout():
    assert parallel.shard_range(160, 3, 8) == (60, 80)
    with pytest.raises(ValueError):
        parallel.shard_range(161, 0, 8)
    lay = parallel.FlatLayout([5, 8, 1])
    assert lay.offsets.tolist() == [0, 8, 16] and lay.total == 20
    flat = torch.zeros(lay.total)
    ts = [torch.arange(5.0), torch.ones(8), torch.tensor([7.0])]
    lay.pack_(flat, ts, scale=0.5)
    v = lay.views(flat, [t.shape for t in ts])
    assert torch.equal(v[0], torch.arange(5.0) * 0.5) and float(v[2]) == 3.5
