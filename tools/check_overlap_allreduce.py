"""Two ranks (torchrun): the all-reduce that starts inside the captured step (external event after the encoder's FC backward) must
produce the same reduced gradients as the plain all-reduce after the step.  Both trainers see batch A, then batch B; after B their
flat gradient buffers are compared (run-to-run TF32 / atomics noise is ~1e-2; a stale read of A's gradients would be O(1))."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, '.')
from preset_gen_vae_b200 import config as pcfg, synthetic  # noqa: E402
from preset_gen_vae_b200.train import TrainStep  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
B = 32
from preset_gen_vae_b200.data.preset import DexedLearnableLayout  # noqa: E402
idx = DexedLearnableLayout().preset_indexes_helper
mc, tc = pcfg.make_default(minibatch_size=B)
pcfg.apply_dataset_dims(mc, idx)
trainers = [TrainStep(mc, tc, idx, process_group=dist.group.WORLD, seed=0, overlap_allreduce=ov) for ov in (True, False)]
for t in trainers:
    t.lr = 0.0                                      # identical weights throughout: the gradients stay comparable
assert trainers[0]._fc_ready is not None and trainers[1]._fc_ready is None
res = []
for t in trainers:
    for seed in (11 + rank, 23 + rank):             # batch A, then batch B (different per rank)
        audio = synthetic.make_audio(B, 1, seed=seed).cuda()
        v_in = synthetic.make_preset_targets(idx, B, seed=seed).cuda()
        info = synthetic.make_sample_info(B).cuda()
        torch.manual_seed(5)                        # same eps / dropout masks for both trainers
        t.step(audio, v_in, info)
        torch.cuda.synchronize()
    res.append(t.gather_sharded(t.flat_grads).clone())     # (the overlapped step leaves every rank with its reduce-scatter shards)
a, b = res
rel = float((a.double() - b.double()).norm() / b.double().norm())
fc = [float((ta.double() - tb.double()).norm() / tb.double().norm()) for ta, tb in
      ((trainers[0]._direct[i], trainers[1]._direct[i]) for i in sorted(trainers[0]._direct))]
other = torch.empty_like(a)
dist.broadcast(other.copy_(a), src=0)
same = bool(torch.equal(other, a))
print("rank %d: overlap vs plain all-reduce: rel-L2 %.3e (FC slices %s); identical on both ranks: %s" % (rank, rel, fc, same), flush=True)
assert rel < 5e-2 and max(fc) < 5e-2 and same
# a step with lr > 0: the segmented optimizer (FC slices first, then the ranges around them) must move every parameter the way the
# single Adam launch of the plain path does
upd = []
for t in trainers:
    before = t.flat_params.clone()
    t.lr = 1e-4
    audio = synthetic.make_audio(B, 1, seed=31 + rank).cuda()
    v_in = synthetic.make_preset_targets(idx, B, seed=31 + rank).cuda()
    info = synthetic.make_sample_info(B).cuda()
    torch.manual_seed(7)
    t.step(audio, v_in, info)
    torch.cuda.synchronize()
    upd.append((t.flat_params - before).double())
ua, ub = upd
cos = float((ua * ub).sum() / (ua.norm() * ub.norm()))
tr = trainers[0]
ranges = [(lo, lo + n) for lo, n in tr._direct_slots] + [(lo, hi) for lo, hi in
          __import__('preset_gen_vae_b200.parallel', fromlist=['x']).complement_segments(tr.flat_grads.numel(), tr._direct_slots)]
moved = [float((ua[lo:hi] != 0).double().mean()) for lo, hi in ranges]
moved_ref = [float((ub[lo:hi] != 0).double().mean()) for lo, hi in ranges]
print("rank %d: parameter update overlap vs plain: cosine %.4f; fraction of elements moved per range %s (plain %s)" %
      (rank, cos, ['%.3f' % m for m in moved], ['%.3f' % m for m in moved_ref]), flush=True)
assert cos > 0.9 and all(abs(a - b) < 0.02 for a, b in zip(moved, moved_ref))
dist.destroy_process_group()
