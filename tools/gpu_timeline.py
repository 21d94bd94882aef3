"""Timeline of one training step at module granularity: start / end of every autograd program (encoder, flows, decoder, losses) forward
and backward, on the stream it runs on, relative to the start of the step.  The step runs eagerly but behind a long blocker kernel, so the
host has enqueued everything before the GPU starts and the streams execute back to back as in the captured graph.
Usage: python tools/gpu_timeline.py [B] [nofork] [noside]"""
import sys

import torch

sys.path.insert(0, '.')
from preset_gen_vae_b200 import config as pcfg, synthetic  # noqa: E402
from preset_gen_vae_b200.data.preset import DexedLearnableLayout  # noqa: E402
from preset_gen_vae_b200.model import ops, program  # noqa: E402
from preset_gen_vae_b200.train import TrainStep  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 160
ops.use_wgrad_fork = 'nofork' not in sys.argv
h = DexedLearnableLayout().preset_indexes_helper
m, t = pcfg.make_default(minibatch_size=B)
pcfg.apply_dataset_dims(m, h)
audio = (torch.rand(B, 1, 88576, device='cuda') - 0.5)
v = synthetic.make_preset_targets(h, B).cuda()
info = synthetic.make_sample_info(B).cuda()
tr = TrainStep(m, t, h, use_cuda_graph=False, overlap_branches='noside' not in sys.argv)
records = []
orig_fwd, orig_bwd = program._ProgramFn.forward, program._ProgramFn.backward


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def fwd(ctx, prog, *a):
    s = ev()
    out = orig_fwd(ctx, prog, *a)
    records.append((type(prog).__name__, 'fwd', torch.cuda.current_stream().cuda_stream, s, ev()))
    return out


def bwd(ctx, *d):
    name = type(ctx.prog).__name__
    s = ev()
    out = orig_bwd(ctx, *d)
    records.append((name, 'bwd', torch.cuda.current_stream().cuda_stream, s, ev()))
    return out


program._ProgramFn.forward = staticmethod(fwd)
program._ProgramFn.backward = staticmethod(bwd)
for _ in range(3):
    tr.step(audio, v, info)
torch.cuda.synchronize()
records.clear()
blocker = torch.empty(16384, 16384, device='cuda')
for _ in range(8):
    torch.mm(blocker, blocker)
t0 = ev()
tr.step(audio, v, info)
t1 = ev()
torch.cuda.synchronize()
print("step (eager, enqueued behind a blocker): %.3f ms; wgrad fork %s, decoder side stream %s" % (t0.elapsed_time(t1), ops.use_wgrad_fork, tr._side is not None))
streams = {}
print("%-28s %-4s %-7s %9s %9s %9s" % ('program', 'dir', 'stream', 'start ms', 'end ms', 'dur ms'))
for name, ph, st, s, e in sorted(records, key=lambda r: t0.elapsed_time(r[3])):
    sid = streams.setdefault(st, len(streams))
    print("%-28s %-4s s%-6d %9.3f %9.3f %9.3f" % (name, ph, sid, t0.elapsed_time(s), t0.elapsed_time(e), s.elapsed_time(e)))
