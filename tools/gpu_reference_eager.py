"""Second reference point (BASELINE.md "not required, but the honest bar"): the reference's own modules - through the oracle port,
plain torch.nn on cuDNN / cuBLAS - running the training-step body of train.py:204-248 in PyTorch eager mode on ONE B200, from a
spectrogram batch that is already on the device (the reference computes its spectrograms in CPU DataLoader workers), with TF32
off and on.  Test infrastructure: nothing here is on the product path.  NOT YET RUN (written after the round's GPU budget was spent):

    gpurun --timeout 600 -- 'python tools/gpu_reference_eager.py 160'
"""
import sys
import time

import torch

sys.path.insert(0, '.')
from oracle import losses as oloss, model as omodel  # noqa: E402
from preset_gen_vae_b200 import config as pcfg, synthetic  # noqa: E402
from preset_gen_vae_b200.data.preset import DexedLearnableLayout  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 160
dev = torch.device('cuda')
helper = DexedLearnableLayout().preset_indexes_helper
m_cfg, t_cfg = pcfg.make_default(minibatch_size=B)
pcfg.apply_dataset_dims(m_cfg, helper)
x = synthetic.make_spectrogram_like(B, 1, seed=0).to(dev)
v_in = synthetic.make_preset_targets(helper, B, seed=0).to(dev)
info = synthetic.make_sample_info(B).to(dev)
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    ext = omodel.build_extended_ae_model(m_cfg, t_cfg, helper)[3].to(dev).train()
    opt = torch.optim.Adam(ext.parameters(), lr=t_cfg.initial_learning_rate, weight_decay=t_cfg.weight_decay, betas=t_cfg.adam_betas)

    def step():
        opt.zero_grad()
        _, _, total = oloss.train_step_losses(ext, x, v_in, info, None, beta=t_cfg.beta)
        total.backward()
        opt.step()
        return total

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    n = 20
    for _ in range(n):
        last = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("reference modules, PyTorch eager on one GPU, TF32 %s: %.2f ms/step = %.0f samples/s (wall %.2f ms/step), loss %.5f"
          % ('on' if tf32 else 'off', ms, B / ms * 1e3, (time.perf_counter() - t0) / n * 1e3, float(last)))
