"""Sweep of an integer debug knob of the kernel library (pgv_debug_set_<knob>): captured training step time at B = 160 per value.
Usage: gpu_sweep_knob.py bn_rows_per_lane 8 16 32"""
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
knob = sys.argv[1]                      # name of a pgv_debug_set_* entry point taking one int, e.g. bn_rows_per_lane, flow_clusters
for rows in [int(a) for a in sys.argv[2:]]:
    getattr(_lib.lib(), 'pgv_debug_set_' + knob)(rows)
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
    print("%s = %3d: %.3f ms per step" % (knob, rows, best), flush=True)
    del tr
    torch.cuda.empty_cache()
