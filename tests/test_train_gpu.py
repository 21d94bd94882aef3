"""GPU: the TrainStep driver (front end -> model -> losses -> backward -> flat-buffer pack -> fused Adam), eager and CUDA-graph."""
import copy

import pytest
import torch

from preset_gen_vae_b200 import config as pcfg, synthetic
from preset_gen_vae_b200.model import loss as ploss, ops
from preset_gen_vae_b200 import train as train_mod
from preset_gen_vae_b200.train import TrainStep

pytestmark = pytest.mark.gpu


def make(B, graph, idx_helper, seed=0):
    m, t = pcfg.make_default(minibatch_size=B)
    pcfg.apply_dataset_dims(m, idx_helper)
    tr = TrainStep(m, t, idx_helper, use_cuda_graph=graph, seed=seed)
    audio = synthetic.make_audio(B, 1, seed=3).cuda()
    v_in = synthetic.make_preset_targets(idx_helper, B, seed=3).cuda()
    info = synthetic.make_sample_info(B).cuda()
    return tr, audio, v_in, info


def test_eager_step_matches_manual_forward_backward_and_torch_adam(idx_helper):
    B = 16
    tr, audio, v_in, info = make(B, False, idx_helper)
    assert tr.flat_params.numel() >= 60372037 and all(p.data_ptr() >= tr.flat_params.data_ptr() for p in tr.params)
    ref_model = copy.deepcopy(tr.model)
    ref_model.ae_model.encoder.fc_weight_grad_out = ref_model.ae_model.decoder.fc_weight_grad_out = None   # plain autograd path
    for m in ref_model.modules():                                                                           # (deepcopy drops the python attribute anyway)
        if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
            assert not hasattr(m.weight, '_pgv_grad_out')
    # the two FC weights and all 16 convolution weights are written straight into the flat gradient buffer
    assert len(tr._direct) == 18 and sum(v.numel() for v in tr._direct.values()) == 1220 * 24576 + 24576 * 610 + 7688592
    ref_opt = torch.optim.Adam(ref_model.parameters(), lr=tr.tc.initial_learning_rate, weight_decay=tr.tc.weight_decay,
                               betas=tr.tc.adam_betas)
    torch.manual_seed(11)
    losses = tr.step(audio, v_in, info)
    # the same step by hand, same RNG stream: module API + torch's Adam
    torch.manual_seed(11)
    x = tr.frontend.compute(audio.view(B, -1), normalize=(-120.0, 0.0)).view(B, 1, 257, 347)
    ref_model.train()
    z0_ml, z0, zk, ld, x_out = ref_model(x, info)
    v_out = ref_model.reg_model(zk)
    rec = ploss.MSELoss()(x_out, x)
    lat = ref_model.latent_loss(z0_ml, z0, zk, ld)
    con = ploss.SynthParamsLoss(idx_helper, True, cat_bce=False, cat_softmax=True, cat_softmax_t=0.2)(v_out, v_in)
    (rec + tr.beta * lat + con).backward()
    got = losses.tolist()
    # monitoring metrics and NaN mask of the step (train.py:232-233, 245) ride along in tr.scalars
    sc = dict(zip(train_mod.SCALAR_NAMES, tr.scalars.tolist()))
    assert abs(sc['controls_qloss'] - ploss.QuantizedNumericalParamsLoss(idx_helper)(v_out, v_in).item()) < 1e-5
    assert abs(sc['controls_accuracy'] - ploss.CategoricalParamsAccuracy(idx_helper)(v_out, v_in).item()) < 0.5
    assert sc['nan_mask'] == 0.0 and sc['flow_input'] == 0.0
    for a, b in zip(got, (rec.item(), lat.item(), con.item())):
        assert abs(a - b) <= 2e-6 * abs(b) + 1e-7
    # Gradients landed in the flat buffer.  Both paths run the same kernels and every reduction is order-fixed since round 2
    # (workspace split-K, finish kernels, per-warp partials), so they agree to rounding: tools/gpu_determinism.py
    # measures 7e-9 run to run (round 1, with fp32 atomics: 4e-2 on the decoder, because this network amplifies 1e-6 activation
    # noise into percent-level gradient differences).  Precision is tested against the fp64 oracle in test_model_gpu.py.
    num = den = dot = 0.0
    for p, q in zip(tr.params, ref_model.parameters()):
        num += float(((p.grad - q.grad) ** 2).sum()); den += float((q.grad ** 2).sum()); dot += float((p.grad * q.grad).sum())
    assert num ** 0.5 <= 1e-4 * den ** 0.5
    for p, q in list(zip(tr.params, ref_model.parameters()))[::17]:
        assert float((p.grad - q.grad).norm()) <= 1e-3 * float(q.grad.norm()) + 1e-7
    # the fused Adam on the flat buffers against torch.optim.Adam fed with the SAME gradients
    for p, q in zip(tr.params, ref_model.parameters()):
        q.grad = p.grad.clone()
    ref_opt.step()
    worst = max(float((p.data - q.data).abs().max()) for p, q in zip(tr.params, ref_model.parameters()))
    assert worst < 5e-6                                   # one Adam step moves weights by <= lr = 2e-4


def test_cuda_graph_steps_train(idx_helper):
    B = 8
    tr, audio, v_in, info = make(B, True, idx_helper)
    before = tr.flat_params.clone()
    hist = []
    for _ in range(6):
        hist.append(tr.step(audio, v_in, info).clone())
    torch.cuda.synchronize()
    hist = torch.stack(hist).cpu()
    assert torch.isfinite(hist).all()
    assert tr._graph is not None and tr.launches_per_step > 150
    assert not torch.equal(before, tr.flat_params)         # every replay applies an optimizer step
    assert tr.step_count == 6
    assert hist[-1, 0] < hist[0, 0] and hist[-1, 2] < hist[0, 2]      # reconstruction and controls losses go down on a fixed batch
    # batched inference tail: audio -> preset parameters in [0, 1]
    v = tr.infer(audio)
    assert v.shape == (B, 610) and float(v.min()) >= 0.0 and float(v.max()) <= 1.0 and tr.model.training
    # host-fed API with input prefetch: same batch from pinned host memory gives the same kind of step
    host = tuple(t.cpu().pin_memory() for t in (audio, v_in, info))
    tr.prefetch(*host)
    l_host = tr.step_prefetched()
    handle = tr.losses_to_host_async()                      # non-blocking read-back of this step's losses ...
    tr.prefetch(*host)
    l_next = tr.step_prefetched().clone()                   # ... while the next step (which overwrites the static loss tensor) runs
    got = handle.get()
    torch.cuda.synchronize()
    assert not got.is_cuda and torch.isfinite(got).all() and float(got[0]) < float(hist[0, 0]) and tr.step_count == 8
    assert not torch.equal(got, l_next.cpu())                # the handle kept step 7's values, not step 8's
    # schedules reach the captured graph through device memory: lr = 0 must freeze the weights
    tr.lr = 0.0
    snap = tr.flat_params.clone()
    tr.step(audio, v_in, info)
    torch.cuda.synchronize()
    assert torch.equal(snap, tr.flat_params)


def test_total_loss_reads_beta_from_device():
    """train.py:228 with the beta warm-up value living in device memory (so a captured graph sees schedule updates)."""
    r, l, c = (torch.tensor(v, device='cuda', requires_grad=True) for v in (1.5, 2.0, 0.25))
    beta = torch.tensor([0.2], device='cuda')
    t = ploss.total_loss(r, l, c, beta)
    t.backward()
    assert abs(t.item() - 2.15) < 1e-6
    assert abs(r.grad.item() - 1.0) < 1e-7 and abs(l.grad.item() - 0.2) < 1e-7 and abs(c.grad.item() - 1.0) < 1e-7
    beta.fill_(0.5)
    assert abs(ploss.total_loss(r, l, c, beta).item() - 2.75) < 1e-6


def test_external_event_recorded_inside_a_graph_orders_later_stream_work():
    """TrainStep's all-reduce overlap (several ranks) relies on this: an event created with external=True and recorded by a node
    INSIDE a captured graph orders work that another stream enqueues after each replay - on every replay, not only the first
    (a stale completion from the previous replay would let the copy below run before this replay's increment)."""
    a = torch.zeros(1 << 20, device='cuda')
    b = torch.zeros_like(a)
    ev = torch.cuda.Event(external=True)
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        torch.cuda._sleep(20_000_000)           # ~10 ms of device time in front of the increment
        a.add_(1)
        ev.record()
        torch.cuda._sleep(2_000_000)
    torch.cuda.synchronize()
    for k in range(1, 5):
        g.replay()
        side.wait_event(ev)
        with torch.cuda.stream(side):
            b.copy_(a, non_blocking=True)
        torch.cuda.synchronize()
        assert float(b[0]) == k and float(b[-1]) == k


def test_dkl_regulariser_nan_guard_and_optimizer_state(idx_helper):
    """latent_flow_input_regularization = 'dkl' (train.py:236-239): no BatchNorm on the encoder output (build.py:24) and
    0.1 * beta * GaussianDkl(mu, logvar) in the loss; NaN guard (train.py:245) as a device flag surfaced by the loss read-back;
    optimizer state save / restore."""
    B = 4
    m, t = pcfg.make_default(minibatch_size=B)
    pcfg.apply_dataset_dims(m, idx_helper)
    t.latent_flow_input_regularization = 'dkl'
    tr = TrainStep(m, t, idx_helper, use_cuda_graph=False, seed=0)
    assert not hasattr(tr.model.ae_model.encoder.mlp, 'lat_in_regularization') or tr.model.ae_model.encoder.out_bn is None
    audio = synthetic.make_audio(B, 1, seed=3).cuda()
    v_in = synthetic.make_preset_targets(idx_helper, B, seed=3).cuda()
    info = synthetic.make_sample_info(B).cuda()
    w0 = tr.flat_params.clone()
    tr.step(audio, v_in, info)
    sc = tr.losses_to_host_async().scalars()
    assert sc['flow_input'] > 0.0 and sc['nan_mask'] == 0.0
    # gradient of the extra term reaches the encoder: same step without it gives different encoder-FC gradients
    g_with = tr.flat_grads.clone()
    state = tr.state_dict()
    assert state['optimizer_state_dict']['step'] == 1 and float(state['optimizer_state_dict']['exp_avg'].abs().sum()) > 0
    tr2 = TrainStep(m, t, idx_helper, use_cuda_graph=False, seed=0)
    tr2.load_state_dict(state)
    assert tr2.step_count == 1 and torch.equal(tr2.exp_avg, tr.exp_avg) and torch.equal(tr2.flat_params, tr.flat_params)
    assert not torch.equal(w0, tr.flat_params)
    tr2.flow_input_dkl = None
    torch.manual_seed(5)
    tr2.step(audio, v_in, info)
    tr.flat_params.copy_(tr2.flat_params)          # irrelevant for the gradient comparison below; keeps both models equal
    assert float((g_with - tr2.flat_grads).abs().max()) > 0.0
    # NaN guard
    tr.model.ae_model.decoder.mlp[0].bias.data[3] = float('nan')        # (a NaN in the audio would be swallowed by the dB floor, like np.maximum)
    tr.step(audio, v_in, info)
    handle = tr.losses_to_host_async()
    with pytest.raises(train_mod.ModelConvergenceError):
        handle.get()
    assert int(handle.scalars()['nan_mask']) & 1                          # the reconstruction loss is the NaN one


def test_pipelined_front_end_is_the_same_training(idx_helper):
    """pipeline_frontend=True: step(batch i) runs the model step of batch i-1 with the front end of batch i on a side branch inside the
    same captured graph.  Same seeds => the same losses, one call later, and the same parameters after the last flush."""
    B = 8
    m, t = pcfg.make_default(minibatch_size=B)
    pcfg.apply_dataset_dims(m, idx_helper)
    plain = TrainStep(m, t, idx_helper, use_cuda_graph=True, seed=0)
    piped = TrainStep(m, t, idx_helper, use_cuda_graph=True, seed=0, pipeline_frontend=True)
    assert piped.pipeline_frontend and torch.equal(plain.flat_params, piped.flat_params)
    batches = [(synthetic.make_audio(B, 1, seed=40 + i).cuda(), synthetic.make_preset_targets(idx_helper, B, seed=40 + i).cuda(),
                synthetic.make_sample_info(B).cuda()) for i in range(3)]
    init = piped.flat_params.clone()
    plain.step(*batches[0])                                  # captures the graph (and trains one step): rewind to the initial state
    torch.cuda.synchronize()
    plain.flat_params.copy_(init); plain.exp_avg.zero_(); plain.exp_avg_sq.zero_(); plain.step_count = 0
    want = []
    for i, b in enumerate(batches):
        torch.cuda.manual_seed(1000 + i)
        want.append(plain.step(*b).clone())
    torch.cuda.synchronize()
    got = []
    assert piped.step(*batches[0]) is None                  # staged only
    for i in range(3):
        torch.cuda.manual_seed(1000 + i)
        out = piped.step(*batches[i + 1]) if i + 1 < 3 else piped.flush_pipeline()
        got.append(out.clone())
    torch.cuda.synchronize()
    assert piped.step_count == 3
    for w, g in zip(want, got):
        assert torch.isfinite(g).all()
        assert float((w - g).abs().max()) <= 2e-3 * float(w.abs().max()), (w.tolist(), g.tolist())
