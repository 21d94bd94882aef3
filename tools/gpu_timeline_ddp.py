"""Timeline of the overlapped data-parallel step (torchrun, >= 2 ranks): when the front end, graph B and the communication stream's
exchange / update phases of three consecutive steps finish, in ms after the first step's start (rank 0)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, '.')
from preset_gen_vae_b200 import config as pcfg, synthetic  # noqa: E402
from preset_gen_vae_b200.data.preset import DexedLearnableLayout  # noqa: E402
from preset_gen_vae_b200.train import TrainStep  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
B = 160
idx = DexedLearnableLayout().preset_indexes_helper
mc, tc = pcfg.make_default(minibatch_size=B)
pcfg.apply_dataset_dims(mc, idx)
t = TrainStep(mc, tc, idx, process_group=dist.group.WORLD, seed=0)
audio = synthetic.make_audio(B, 1, seed=3 + rank).cuda()
v_in = synthetic.make_preset_targets(idx, B, seed=3 + rank).cuda()
info = synthetic.make_sample_info(B).cuda()
for _ in range(6):
    t.step(audio, v_in, info)
torch.cuda.synchronize()
dist.barrier()
t._trace = []
for _ in range(3):
    t.step(audio, v_in, info)
torch.cuda.synchronize()
if rank == 0:
    t0 = t._trace[0][1]
    print("phase segments %s; fractions of the gradient bytes %s" % (t._phase_segments, ['%.3f' % f for f in t.phase_fractions]))
    for name, ev in sorted(t._trace, key=lambda ne: t0.elapsed_time(ne[1])):
        print("  %8.3f ms  %s" % (t0.elapsed_time(ev), name))
dist.destroy_process_group()
