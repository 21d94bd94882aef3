"""Run-to-run reproducibility of one eager training step: same parameters, same RNG seed, twice; per-tensor relative
difference of the gradients (sorted).  Separates real nondeterminism (atomics order, races) from test tolerances."""
import sys

import torch

sys.path.insert(0, '.')
from preset_gen_vae_b200 import config as pcfg, synthetic  # noqa: E402
from preset_gen_vae_b200.data.preset import DexedLearnableLayout  # noqa: E402
from preset_gen_vae_b200.model import ops  # noqa: E402
from preset_gen_vae_b200.train import TrainStep  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
if len(sys.argv) > 2:
    ops.set_precision(sys.argv[2])
h = DexedLearnableLayout().preset_indexes_helper
m, t = pcfg.make_default(minibatch_size=B)
pcfg.apply_dataset_dims(m, h)
tr = TrainStep(m, t, h, use_cuda_graph=False)
audio = synthetic.make_audio(B, 1, seed=3).cuda()
v = synthetic.make_preset_targets(h, B, seed=3).cuda()
info = synthetic.make_sample_info(B).cuda()
names = [n for n, p in tr.model.named_parameters() if p.requires_grad]
runs = []
for r in range(3):
    torch.manual_seed(11)
    tr._refresh_hyper()
    losses = tr._device_step(audio, v, info, with_optimizer=False)
    torch.cuda.synchronize()
    runs.append((losses.tolist(), [p.grad.clone() for p in tr.params]))
print('precision', ops.get_precision(), 'B', B)
for r in (1, 2):
    print('run %d vs run 0: losses' % r, runs[0][0], runs[r][0])
    rows = []
    for n, a, b in zip(names, runs[0][1], runs[r][1]):
        d = float((a - b).norm()); nrm = float(a.norm())
        rows.append((d / max(nrm, 1e-30), n, nrm))
    tot2 = sum((rel * nrm) ** 2 for rel, n, nrm in rows)
    rows.sort(key=lambda r: -(r[0] * r[2]))
    for rel, n, nrm in rows[:12]:
        print('   rel %.3e  share of the squared difference %.2f  %-60s |g| = %.3e' % (rel, (rel * nrm) ** 2 / max(tot2, 1e-300), n, nrm))
    tot = torch.sqrt(sum(((a - b) ** 2).sum() for a, b in zip(runs[0][1], runs[r][1]))) / torch.sqrt(sum((a ** 2).sum() for a in runs[0][1]))
    print('   global rel-L2 %.3e' % float(tot))
