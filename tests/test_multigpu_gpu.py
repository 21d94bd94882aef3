"""GPU, world size 2 (skipped on a one-GPU box): the overlapped data-parallel step - two captured graphs, FC gradient slices reduced
and updated under the encoder backward, the tail under the next front end - must give the same reduced gradients and the same
parameter update as the plain path (all-reduce after the step, one Adam launch), identically on both ranks."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_overlapped_data_parallel_step_matches_the_plain_one():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29731', os.path.join(ROOT, 'tools', 'check_overlap_allreduce.py')]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0


def test_gathered_batch_normaliser_of_the_controls_loss(idx_helper):
    """SynthParamsLoss divides every categorical group's cross-entropy by its number of useful rows in the batch (loss.py:172) and the
    reference's DataParallel evaluates it on the gathered batch.  With the all-reduced counts / world as `group_counts`, the mean of the
    per-shard losses (what data-parallel ranks produce) equals the full-batch loss, and so do the gradients."""
    from preset_gen_vae_b200 import synthetic
    from preset_gen_vae_b200.model import loss as ploss
    B, W = 48, 2
    v_in = synthetic.make_preset_targets(idx_helper, B, seed=9, p_silent_operator=0.3).cuda()
    g = torch.Generator(device='cuda').manual_seed(1)
    v_out = torch.rand(B, 610, device='cuda', generator=g).requires_grad_()
    crit = ploss.SynthParamsLoss(idx_helper, True, cat_bce=False, cat_softmax=True, cat_softmax_t=0.2)
    full = crit(v_out, v_in)
    full.backward()
    g_full = v_out.grad.clone()
    v_out.grad = None
    counts = crit.useful_counts(v_in)
    assert counts.dtype == torch.float64 and float(counts.max()) <= B
    shard_losses = []
    for r in range(W):
        sl = slice(r * B // W, (r + 1) * B // W)
        assert not torch.equal(crit.useful_counts(v_in[sl]), counts / W)          # the shards do differ in their own counts
        l = crit(v_out[sl], v_in[sl], group_counts=counts / W)
        (l / W).backward()
        shard_losses.append(l)
    mean = sum(shard_losses) / W
    assert abs(mean.item() - full.item()) < 2e-6 * abs(full.item())
    assert float((v_out.grad - g_full).abs().max()) < 1e-6 * float(g_full.abs().max()) + 1e-9
