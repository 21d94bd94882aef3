"""Where the training step's device time goes: every C entry point of one eager, single-stream step bracketed by CUDA events (GPU kept
busy so that launch gaps are not counted), next to the captured-graph step time.  Usage: python tools/gpu_step_breakdown.py [B] [c6]"""
import sys

import torch

sys.path.insert(0, '.')
from preset_gen_vae_b200 import config as pcfg, synthetic  # noqa: E402
from preset_gen_vae_b200.data.preset import DexedLearnableLayout  # noqa: E402
from preset_gen_vae_b200.model import ops  # noqa: E402
from preset_gen_vae_b200.train import TrainStep  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 160
c6 = 'c6' in sys.argv
h = DexedLearnableLayout().preset_indexes_helper
six = ((40, 85), (50, 85), (60, 42), (60, 85), (60, 127), (70, 85))
m, t = pcfg.make_default(minibatch_size=B, **(dict(midi_notes=six, stack_spectrograms=True) if c6 else {}))
pcfg.apply_dataset_dims(m, h)
C = m.input_tensor_size[1]
audio = (torch.rand(B, C, 88576, device='cuda') - 0.5)
v = synthetic.make_preset_targets(h, B).cuda()
info = synthetic.make_sample_info(B).cuda()


def time_steps(tr, n=10):
    for _ in range(4):
        tr.step(audio, v, info)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        tr.step(audio, v, info)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


tr = TrainStep(m, t, h, use_cuda_graph=True)
print("captured-graph step: %.3f ms (B=%d, C=%d, %d launches)" % (time_steps(tr), B, C, tr.launches_per_step))
for fork, side in ((False, False), (True, True)):
    ops.use_wgrad_fork = fork
    tr2 = TrainStep(m, t, h, use_cuda_graph=True, overlap_branches=side)
    print("  graph step with wgrad fork %s, decoder side stream %s: %.3f ms" % (fork, side, time_steps(tr2)))
    del tr2
ops.use_wgrad_fork = False
tr.use_graph, tr._side = False, None
tr.step(audio, v, info)
torch.cuda.synchronize()
blocker = torch.empty(16384, 16384, device='cuda')
for _ in range(6):
    torch.mm(blocker, blocker)
ops.start_profile()
for _ in range(2):
    tr.step(audio, v, info)
prof = ops.stop_profile()
tot = sum(v_['ms'] for v_ in prof.values()) / 2
print("eager single-stream step, sum over entry points: %.3f ms" % tot)
print("%-34s %5s %9s %7s %9s %9s" % ('entry point', 'calls', 'ms/step', 'share', 'TFLOP/s', 'GB/s'))
for name, v_ in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
    ms = v_['ms'] / 2
    print("%-34s %5d %9.4f %6.1f%% %9.1f %9.0f" % (name, v_['calls'] // 2, ms, 100 * ms / tot, v_['flops'] / 2 / (ms * 1e-3) / 1e12 if v_['flops'] else 0.0,
                                                 v_['bytes'] / 2 / (ms * 1e-3) / 1e9 if v_['bytes'] else 0.0))
