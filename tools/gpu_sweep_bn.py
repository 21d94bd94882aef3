"""Sweep of the channels-last BatchNorm grid sizing knob: captured training step time at B = 160 for several rows-per-thread values."""
import sys

import torch

sys.path.insert(0, '.')
from preset_gen_vae_b200 import _lib, config as pcfg, synthetic  # noqa: E402
from preset_gen_vae_b200.data.preset import DexedLearnableLayout  # noqa: E402
from preset_gen_vae_b200.train import TrainStep  # noqa: E402

B = 160
h = DexedLearnableLayout().preset_indexes_helper
audio = synthetic.make_audio(B, 1, seed=3).cuda()
v = synthetic.make_preset_targets(h, B, seed=3).cuda()
info = synthetic.make_sample_info(B).cuda()
for rows in [int(a) for a in sys.argv[1:]] or [8, 4, 2, 1]:
    _lib.lib().pgv_debug_set_bn_rows_per_lane(rows)
    m, t = pcfg.make_default(minibatch_size=B)
    pcfg.apply_dataset_dims(m, h)
    tr = TrainStep(m, t, h, pipeline_frontend=True)
    for _ in range(8):
        tr.step(audio, v, info)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            tr.step(audio, v, info)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 20)
    print("rows per thread %2d: %.3f ms per step" % (rows, best), flush=True)
    del tr
    torch.cuda.empty_cache()
