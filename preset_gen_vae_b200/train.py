"""One data-parallel training step of the hot path, as the reference's loop body does it (train.py:204-248):

    audio -> mel front end (+ min-max) -> ExtendedAE forward -> regression head -> MSE + beta * latent + controls
          -> backward -> (N > 1: gradient all-reduce, mean over ranks) -> Adam (L2 weight decay)

B200 design: one process per GPU, parameters resident per rank in ONE flat fp32 buffer (module
parameters are views into it), gradients packed into one flat buffer by a single kernel, one NCCL all-reduce over
NVLink on that buffer, one fused Adam kernel over the flat buffers, and the whole device-side step replayed from a
CUDA graph.  BatchNorm statistics are per-rank, as with the reference's nn.DataParallel replicas (train.py:95-97).
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib, parallel, synthetic
from .model import build, loss as ploss, ops
from .utils import audio as paudio

_DEBUG_SKIP_AR = os.environ.get('PGV_DEBUG_SKIP_ALLREDUCE') == '1'      # measurement only (tools/): isolates what the exchange costs the step


class ModelConvergenceError(ValueError):
    """utils/exception.py:9-10 of the reference."""


# layout of TrainStep.scalars (device, fp32): the step's un-weighted losses, the two monitoring metrics of train.py:232-233, the flow-input
# regulariser of train.py:236-239 (0 unless latent_flow_input_regularization == 'dkl') and the NaN bit mask of train.py:245
SCALAR_NAMES = ('recons', 'latent', 'controls', 'controls_qloss', 'controls_accuracy', 'flow_input', 'nan_mask')


def check_nan_mask(mask, step=None):
    """utils/exception.py:13-22: raises ModelConvergenceError naming the first NaN loss tensor."""
    mask = int(mask)
    if mask:
        names = ('recons_loss', 'lat_loss', 'flow_input_loss', 'cont_loss')
        bad = [names[i] for i in range(4) if mask >> i & 1]
        raise ModelConvergenceError("Step {}: {} contain(s) a nan item".format(step, ', '.join(bad)))


class _HostLosses:
    """Pending device -> host copy of a step's scalars; see TrainStep.losses_to_host_async."""

    def __init__(self, buf, event, step):
        self._buf, self._event, self._step = buf, event, step

    def get(self, check_nan=True):
        """(recons, latent, controls) of that step; raises ModelConvergenceError if one of its loss terms was NaN (train.py:245)."""
        self._event.synchronize()
        if check_nan:
            check_nan_mask(self._buf[6].item(), self._step)
        return self._buf[:3].clone()

    def scalars(self):
        self._event.synchronize()
        return {k: float(v) for k, v in zip(SCALAR_NAMES, self._buf.tolist())}


class TrainStep:
    def __init__(self, model_config, train_config, idx_helper, device=None, process_group=None, use_cuda_graph=True,
                 spec_stats=None, beta=None, seed=0, overlap_branches=True, overlap_allreduce=True, pipeline_frontend=False):
        self.mc, self.tc, self.idx_helper = model_config, train_config, idx_helper
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.pg = process_group
        self.world = 1 if process_group is None else torch.distributed.get_world_size(process_group)
        self.use_graph = use_cuda_graph
        # One GPU: software-pipeline the front end.  step(batch i) then runs the model step of batch i-1 and, on a side branch forked
        # where the latent flow's backward starts (300 launches of 5-10 us that leave the SMs idle), the front end of batch i; it
        # returns the losses of batch i-1 (None on the first call).  flush_pipeline() runs the last staged batch.
        self.pipeline_frontend = bool(pipeline_frontend) and process_group is None and use_cuda_graph
        self._pgraph = None
        self.spec_stats = spec_stats or synthetic.SPEC_STATS
        self.beta = train_config.beta if beta is None else beta
        torch.manual_seed(seed)                       # same initial weights on every rank
        with torch.cuda.device(self.device):
            self.model = build.build_extended_ae_model(model_config, train_config, idx_helper)[3].to(self.device)
        self.model.train()
        # ... but independent noise (eps, dropout masks) per rank, like the per-device generators of the reference's DataParallel replicas
        self.rank = 0 if process_group is None else torch.distributed.get_rank(process_group)
        with torch.cuda.device(self.device):
            torch.cuda.manual_seed(seed * 1000 + self.rank)
        if not self.model.is_flow_based_latent_space:
            raise NotImplementedError("TrainStep drives FlowVAE models (BasicVAE cannot be reached through ExtendedAE, SURVEY.md 9.10)")
        reg = train_config.latent_flow_input_regularization.lower()
        if reg not in ('bn', 'dkl', 'none'):
            raise NotImplementedError("latent_flow_input_regularization = %r" % train_config.latent_flow_input_regularization)
        self.flow_input_dkl = ploss.GaussianDkl(normalize=train_config.normalize_losses) if reg == 'dkl' else None     # train.py:128, 236-239
        if not model_config.forward_controls_loss:
            raise NotImplementedError("TrainStep implements the forward controls loss (config default); FlowParamsLoss (model/loss.py) is "
                                      "available as a criterion for custom loops")
        self.metrics_criterion = ploss.PresetMetrics(idx_helper)                                                   # train.py:121-124
        # decoder branch on a side stream, concurrent with the regression-flow branch (both only depend on z_K)
        self._side = torch.cuda.Stream(device=self.device) if overlap_branches else None
        self.frontend = paudio.build_spectrogram(model_config, device=self.device)
        self.recons_criterion = ploss.MSELoss() if train_config.normalize_losses else ploss.L2Loss()
        self.controls_criterion = ploss.SynthParamsLoss(
            idx_helper, train_config.normalize_losses, cat_bce=train_config.params_cat_bceloss,
            cat_softmax=(not model_config.params_reg_softmax and not train_config.params_cat_bceloss),
            cat_softmax_t=train_config.params_cat_softmax_temperature)
        self._flatten_parameters()
        # Several ranks: the two FC weight gradients (three quarters of the gradient bytes) are final as soon as the encoder's
        # backward has passed its FC layer, long before the captured step ends.  An EXTERNAL event recorded at that point inside
        # the graph lets a communication stream start their all-reduce under the encoder's convolution backward.
        self._fc_ready = None
        if self.world > 1 and use_cuda_graph and overlap_allreduce:
            self._fc_ready = torch.cuda.Event(external=True)
            self._dec_ready = torch.cuda.Event(external=True)
            self._comm_stream = torch.cuda.Stream(device=self.device)
            self._aux_stream = torch.cuda.Stream(device=self.device)
        self.step_count = 0
        self.lr = train_config.initial_learning_rate
        # lr, bias corrections, grad scale, beta: staged through a ring of pinned buffers because the host runs ahead of the device
        # (an asynchronous copy reads its pinned source when it EXECUTES, so a buffer is only rewritten after its last copy is done)
        self._hyper_ring = [torch.zeros(5, dtype=torch.float32).pin_memory() for _ in range(8)]
        self._hyper_events = [None] * len(self._hyper_ring)
        self._hyper_dev = torch.zeros(5, dtype=torch.float32, device=self.device)
        self._graph = None
        self._static = None
        self.losses = None
        self.scalars = None
        self._prepared_params = []
        self._persistent_operands = False
        self._update_done = None
        self._nan_mask = torch.zeros(1, dtype=torch.int32, device=self.device)
        # Data parallel: SynthParamsLoss normalises every categorical group by its number of useful rows in the batch (loss.py:172) and the
        # reference evaluates it on the GATHERED batch.  The counts only depend on the targets: each step all-reduces the 54 counts
        # before the (captured) step and the criterion divides by (global count / world), so the mean of the per-rank losses - and of
        # their gradients - is the gathered-batch value.
        self._group_counts = None
        if self.world > 1:
            self._group_counts = torch.zeros(self.controls_criterion._tables.n_grp, dtype=torch.float64, device=self.device)

    # ------------------------------------------------------------------ flat parameter / gradient / Adam-state buffers
    def _flatten_parameters(self):
        params = [p for p in self.model.parameters() if p.requires_grad]
        sizes = [p.numel() for p in params]
        # 16-byte aligned slots; with several ranks every slot is a multiple of 4 x world elements, so that any run of whole slots
        # splits into `world` equal, 16-byte aligned shards (reduce-scatter / sharded Adam / all-gather in _step_overlapped)
        self.layout = parallel.FlatLayout(sizes, align_elems=4 * max(1, self.world))
        flat = torch.zeros(self.layout.total, dtype=torch.float32, device=self.device)
        for p, view in zip(params, self.layout.views(flat, [p.shape for p in params])):
            view.copy_(p.data)
            p.data = view
        self.params, self.flat_params, self._offs, self._sizes = params, flat, self.layout.offsets, sizes
        self.flat_grads = torch.zeros_like(flat)
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.n_param_elems = sum(sizes)
        # Weight-gradient kernels write straight into the flat buffer wherever they can produce the parameter's own layout: the two
        # fully-connected weights (three quarters of all gradient bytes) and every convolution weight (the split-K finish pass of the
        # channels-last kernel emits the PyTorch layout).  Gradients are stored unscaled; with several ranks the all-reduce sums them and
        # the 1/world factor is the grad_scale the fused Adam kernel reads from device memory.
        self._direct = {}
        views = self.layout.views(self.flat_grads, [p.shape for p in params])
        index_of = {id(p): i for i, p in enumerate(params)}
        enc, dec = self.model.ae_model.encoder, self.model.ae_model.decoder
        for owner, lin in ((enc, enc.mlp[1]), (dec, dec.mlp[0])):
            idx = index_of[id(lin.weight)]
            owner.fc_weight_grad_out = views[idx]
            self._direct[idx] = views[idx]
        self._fc_slots = [(int(self._offs[i]), sizes[i]) for i in sorted(self._direct)]
        conv_weights = [enc.features_mixer_cnn, dec.features_unmixer_cnn]
        if enc.spectrogram_channels == 1:              # the per-channel CNNs are shared over the channels: their gradients are summed
            conv_weights += [enc.single_ch_cnn, dec.single_ch_cnn]
        for owner in conv_weights:
            for m in owner.modules():
                if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
                    idx = index_of[id(m.weight)]
                    m.weight._pgv_grad_out = views[idx]
                    self._direct[idx] = views[idx]
        # each flow's parameters are one contiguous slice of the flat buffer: registered for the L2 prefetch ahead of its kernel chain
        for flow in (self.model.ae_model.flow_transform, getattr(self.model.reg_model, '_forward_flow_transform', None)):
            if flow is not None:
                idxs = [index_of[id(p)] for p in flow.parameters()]
                lo, hi = min(idxs), max(idxs)
                if hi - lo + 1 == len(idxs):
                    flow._pgv_param_range = flat[int(self._offs[lo]):int(self._offs[hi]) + sizes[hi]]
        self._packed = [i for i in range(len(params)) if i not in self._direct]
        self._direct_slots = self._fc_slots                                             # (offset, size) of the early all-reduce slices
        enc_ids = {id(p) for p in enc.parameters()}
        self._in_encoder = [id(p) in enc_ids for p in params]
        self._pack_tables = {}                          # key -> (indices, pinned host table, device table)
        self._early_packed = []                         # per phase: indices packed by _pack_early in this step (overlapped data-parallel step)
        self.early_adam = False                         # one GPU: Adam on everything outside the encoder stack under the encoder backward

    def _pack_subset(self, key, idx, scale=1.0):
        """Copies the gradients of parameters `idx` into their flat slots (one launch; the pointer table is re-read from pinned host
        memory by the captured copy node, and is the same every step because the captured allocations are)."""
        if not idx:
            return
        tab = self._pack_tables.get(key)
        if tab is None or tab[0] != idx:
            host = torch.zeros(len(idx) * 3, dtype=torch.int64).pin_memory()
            host[1::3] = torch.from_numpy(np.asarray([self._offs[i] for i in idx], dtype=np.int64))
            host[2::3] = torch.tensor([self._sizes[i] for i in idx], dtype=torch.int64)
            tab = self._pack_tables[key] = (list(idx), host, torch.zeros(len(idx) * 3, dtype=torch.int64, device=self.device))
        _, host, dev = tab[:3]
        capturing = torch.cuda.is_current_stream_capturing()
        if len(tab) > 3 and tab[3] is not None and not capturing:
            tab[3].synchronize()                     # eager mode: the host runs ahead; the previous copy must have read the table
        host[0::3] = torch.tensor([self.params[i].grad.data_ptr() for i in idx], dtype=torch.int64)
        dev.copy_(host, non_blocking=True)
        if not capturing:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._pack_tables[key] = tab[:3] + (ev,)
        _lib.check(_lib.lib().pgv_multi_pack(_lib.ptr(dev), len(idx), max(self._sizes[i] for i in idx),
                                             _lib.ptr(self.flat_grads), float(scale), _lib.stream_ptr(self.device)), 'pgv_multi_pack')
        ops.launches += 1

    def _pack_early(self, phase):
        """Overlapped data-parallel step: packs the gradients that are final at one of two points of the backward pass, so that their
        exchange + Adam run under the rest of it instead of after the step.
          phase 0  from the latent flow's backward start: the decoder, the regression flow and everything else whose backward node
                   has already run and accumulated;
          phase 1  from the encoder's backward right before it records `_fc_ready`: what became final since (the latent flow)."""
        done = {i for ph in self._early_packed for i in ph}
        idx = [i for i in self._packed if i not in done and not self._in_encoder[i] and self.params[i].grad is not None]
        self._pack_subset('early%d' % phase, idx)
        while len(self._early_packed) <= phase:
            self._early_packed.append([])
        self._early_packed[phase] = idx

    def _early_update(self):
        """One GPU: called from the encoder's backward once it has passed its FC layer.  Every gradient outside the encoder's
        convolution stack is final (94 % of the parameters): they are packed and Adam runs on their slots on a side stream, under the
        encoder's convolution backward, instead of 0.3 ms at the end of the step (nothing that is still to run reads those
        parameters)."""
        main = torch.cuda.current_stream(self.device)
        self._pack_early(0)
        slot_index = {(int(self._offs[i]), self._sizes[i]): i for i in self._direct}
        enc_fc = {slot_index[sl] for sl in self._fc_slots if self._in_encoder[slot_index[sl]]}
        early = set(self._early_packed[0]) | enc_fc | {i for i in self._direct if not self._in_encoder[i]}
        segs = parallel.merged_slot_ranges(self._offs, self.flat_grads.numel(), early)
        if getattr(self, '_update_stream', None) is None:
            self._update_stream = torch.cuda.Stream(device=self.device)
        self._update_stream.wait_stream(main)
        with torch.cuda.stream(self._update_stream):
            for lo, hi in segs:
                self._adam(lo, hi)
        self._early_adam_segments = segs

    def _pack_grads(self, scale=1.0):
        for i, view in self._direct.items():           # already in place; expose them like every other gradient
            self.params[i].grad = view
        if self._early_packed:
            done = {i for ph in self._early_packed for i in ph}
            self._pack_subset('late', [i for i in self._packed if i not in done], scale)
        else:
            self._pack_subset('all', self._packed, scale)

    # ------------------------------------------------------------------ operand copies of the weights, off the critical path
    def _prepare_operands(self):
        """The tensor-core kernels consume TF32-rounded, re-laid-out copies of the weights (forward / data-gradient matrices of every
        convolution, 16-byte-pitched FC matrices: 0.3 ms of small copy kernels per step).  They only depend on the parameters, so they are
        enqueued on a side stream at the very start of the step and run under the front end's DFT / mel contractions; the model picks
        them up through `ops.prepared_of(param)`."""
        if self._persistent_operands:
            return                                             # refreshed behind the optimizer instead (_refresh_operands)
        main = torch.cuda.current_stream(self.device)
        if getattr(self, '_prep_stream', None) is None:
            self._prep_stream = torch.cuda.Stream(device=self.device)
        self._prep_stream.wait_stream(main)
        self._prepared_params = []
        with torch.cuda.stream(self._prep_stream):
            self._refresh_operands(conv=True, fc=True, persistent=False)

    def _operand_modules(self):
        convs = []
        for m in self.model.modules():
            if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
                cout, cin, kh, kw = m.weight.shape            # (ConvTranspose2d: the convolution it is the data gradient of)
                if ops.cl_mode() and _lib.lib().pgv_conv_cl_supported(cin, cout, kh, kw, m.stride[0], m.padding[0]):
                    convs.append(m)
        enc, dec = self.model.ae_model.encoder, self.model.ae_model.decoder
        fcs = [lin for lin in (enc.mlp[1], dec.mlp[0]) if ops.fc_route(self.tc.minibatch_size, *lin.weight.shape) == 'cl']
        return convs, fcs

    def _refresh_operands(self, conv, fc, persistent):
        """(Re)computes the operand copies on the current stream; persistent: in place, into the buffers the captured graph reads."""
        convs, fcs = self._operand_modules()
        cur = torch.cuda.current_stream(self.device)

        def done(w):
            if not persistent:                                 # consumers wait for THIS parameter's copies only (ops.prepared_of)
                w._pgv_prepared_event = torch.cuda.Event()
                w._pgv_prepared_event.record(cur)
                self._prepared_params.append(w)
        enc_params = {id(p) for p in self.model.ae_model.encoder.parameters()}
        if conv in ('enc', 'dec'):                             # one side's convolutions only
            convs = [m for m in convs if (id(m.weight) in enc_params) == (conv == 'enc')]
        if fc in ('enc', 'dec'):
            fcs = [lin for lin in fcs if (id(lin.weight) in enc_params) == (fc == 'enc')]
        jobs = [(m.weight, m) for m in convs] * bool(conv) + [(lin.weight, None) for lin in fcs] * bool(fc)
        jobs.sort(key=lambda j: (id(j[0]) not in enc_params, j[1] is None))      # encoder first (convolutions, then its FC), then the decoder
        for w, m in jobs:
            if m is not None:
                w._pgv_prepared = ops.prep_conv_weights(w, m.stride[0], m.padding[0], out=w._pgv_prepared if persistent else None)
            else:
                w._pgv_prepared = ops.round_copy(w, (w.shape[1] + 3) // 4 * 4, out=w._pgv_prepared if persistent else None)
            done(w)

    def _make_operands_persistent(self):
        """Several ranks: the copies live in fixed buffers that the model graph reads and that are refreshed on the communication
        stream right behind the Adam launch of their parameters (FC copies under the encoder backward, the rest under the next step's
        front end)."""
        self._drop_prepared()
        self._refresh_operands(conv=True, fc=True, persistent=False)      # allocates; from now on refreshed in place
        torch.cuda.current_stream(self.device).synchronize()
        for w in self._prepared_params:                                   # ordered by the graphs from now on, not by per-parameter events
            w._pgv_prepared_event = None
        self._prepared_params = []
        self._persistent_operands = True

    def _join_prepared(self):
        """End of the step: the side stream has nothing pending that the main stream has not waited for, except copies nobody used."""
        if not self._persistent_operands and getattr(self, '_prep_stream', None) is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._prep_stream)

    def _drop_prepared(self):
        if self._persistent_operands:
            return
        for w in self._prepared_params:
            w._pgv_prepared = None
            w._pgv_prepared_event = None
        self._prepared_params = []

    # ------------------------------------------------------------------ one step, eager (also what gets captured)
    def _device_step(self, audio, v_in, sample_info, with_optimizer):
        self._prepare_operands()                      # side stream: runs under the front end
        try:
            x_in = self._front_end(audio)
            return self._model_step(x_in, v_in, sample_info, with_optimizer)
        finally:
            self._drop_prepared()

    def _front_end(self, audio, out=None):
        B, C, L = audio.shape
        x_in = self.frontend.compute(audio.view(B * C, L), normalize=(self.spec_stats['min'], self.spec_stats['max']), out=out)
        ops.launches += _lib.lib().pgv_frontend_launch_count(self.mc.mel_bins)
        return x_in.view(B, C, x_in.shape[-2], x_in.shape[-1])

    def _model_step(self, x_in, v_in, sample_info, with_optimizer):
        """Everything after the front end: forward, losses, backward, gradient packing (and Adam when with_optimizer)."""
        if not self._prepared_params:                 # (not reached through _device_step)
            self._prepare_operands()
        self.model.ae_model.decoder_stream = self._side
        try:
            z0_ml, z0, zk, logdet, x_out = self.model(x_in, sample_info)
        finally:
            self.model.ae_model.decoder_stream = None
        v_out = self.model.reg_model(zk)
        with torch.no_grad():                        # monitoring metrics, before the criterion like train.py:232-233
            metrics = self.metrics_criterion(v_out, v_in)
        cont = self.controls_criterion(v_out, v_in, group_counts=self._group_counts)
        lat = self.model.latent_loss(z0_ml, z0, zk, logdet)
        flow_in = self.flow_input_dkl(z0_ml[:, 0, :], z0_ml[:, 1, :], packed=z0_ml) if self.flow_input_dkl is not None else None
        if self._side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._side)
        recons = self.recons_criterion(x_out, x_in)
        total = ploss.total_loss(recons, lat, cont, self._hyper_dev[4:5], flow_in, 0.1)
        self._nan_mask.zero_()
        ops.nan_flags_(self._nan_mask, recons, lat, flow_in if flow_in is not None else recons, cont)               # train.py:245
        for p in self.params:
            p.grad = None
        self._early_packed = []
        enc = self.model.ae_model.encoder
        early_adam = with_optimizer and self.world == 1 and self.early_adam and getattr(enc, 'before_fc_grads_ready', None) is None
        self._early_adam_segments = None
        if early_adam:
            enc.before_fc_grads_ready = self._early_update
        try:
            total.backward()
        finally:
            if early_adam:
                enc.before_fc_grads_ready = None
        self._pack_grads(1.0)
        self._join_prepared()
        if with_optimizer:
            if self._early_adam_segments is not None:          # the rest: the encoder's stack (6 % of the parameters)
                n = self.flat_grads.numel()
                for lo, hi in parallel.complement_segments(n, [(lo, hi - lo) for lo, hi in self._early_adam_segments]):
                    self._adam(lo, hi)
                torch.cuda.current_stream(self.device).wait_stream(self._update_stream)
            else:
                self._adam()
        zero = torch.zeros((), device=self.device)
        return torch.stack([recons.detach(), lat.detach(), cont.detach(), metrics[0], metrics[1],
                            flow_in.detach() if flow_in is not None else zero, self._nan_mask[0].float()])

    def _adam(self, lo=0, hi=None):
        """Fused Adam on elements [lo, hi) of the flat buffers, on the current stream."""
        tc = self.tc
        hi = self.flat_params.numel() if hi is None else hi
        sl = slice(lo, hi)
        _lib.check(_lib.lib().pgv_adam_step_dev(
            _lib.ptr(self.flat_params[sl]), _lib.ptr(self.flat_grads[sl]), _lib.ptr(self.exp_avg[sl]), _lib.ptr(self.exp_avg_sq[sl]),
            hi - lo, _lib.ptr(self._hyper_dev), tc.adam_betas[0], tc.adam_betas[1], 1e-8, tc.weight_decay,
            _lib.stream_ptr(self.device)), 'pgv_adam_step_dev')
        ops.launches += 1

    def _refresh_hyper(self):
        self.step_count += 1
        b1, b2 = self.tc.adam_betas
        slot = self.step_count % len(self._hyper_ring)
        host = self._hyper_ring[slot]
        if self._hyper_events[slot] is not None:
            self._hyper_events[slot].synchronize()
        host[0] = self.lr
        host[1] = 1.0 - b1 ** self.step_count
        host[2] = float(np.sqrt(1.0 - b2 ** self.step_count))
        host[3] = 1.0 / self.world                      # grad_scale: the all-reduce SUMS the per-rank gradients
        host[4] = self.beta
        self._hyper_dev.copy_(host, non_blocking=True)
        self._hyper_events[slot] = torch.cuda.Event()
        self._hyper_events[slot].record(torch.cuda.current_stream(self.device))

    def _allreduce(self, overlapped=False):
        """Sum of the flat gradient buffer over the ranks (Adam applies the 1/world factor through grad_scale).  `overlapped`: the
        step that was just launched is the captured graph containing the `_fc_ready` record, so the FC slices are reduced from the
        communication stream as soon as that event fires and only the remaining segments wait for the end of the step."""
        if not overlapped:
            parallel.allreduce_mean_(self.flat_grads, self.pg)
            return
        parallel.allreduce_segments_(self.flat_grads, self._direct_slots, self.pg, early_stream=self._comm_stream, early_event=self._fc_ready)

    def step(self, audio, v_in, sample_info):
        """audio [B, C, L] fp32, v_in [B, L_params] fp32, sample_info [B, 3] int32: CUDA tensors on this rank's device.
        Returns a device tensor (recons, latent, controls) of this rank's un-weighted losses."""
        if self.world > 1 and self.use_graph and self._fc_ready is not None:
            return self._step_overlapped(audio, v_in, sample_info)
        if self.pipeline_frontend:
            return self._step_pipelined(audio, v_in, sample_info)
        self.finish_updates()
        self._refresh_hyper()
        fused_opt = self.world == 1
        if self._group_counts is not None:
            counts = parallel.global_useful_counts(self.controls_criterion.useful_counts(v_in), self.pg)
            self._group_counts.copy_(counts / self.world)
        if not self.use_graph:
            scalars = self._device_step(audio, v_in, sample_info, with_optimizer=fused_opt)
        else:
            if self._graph is None:
                self._capture(audio, v_in, sample_info, fused_opt)
            for dst, src in zip(self._static[:3], (audio, v_in, sample_info)):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
            self._graph.replay()
            scalars = self._static[3].clone()          # the graph's output buffer is overwritten by the next replay
        if not fused_opt:
            self._allreduce(overlapped=self.use_graph and self._fc_ready is not None)
            self._adam()
            if self._persistent_operands:
                self._refresh_operands(conv=True, fc=True, persistent=True)
        self.scalars = scalars                         # SCALAR_NAMES
        self.losses = scalars[:3]
        return self.losses

    # ------------------------------------------------------------------ one GPU: front end of the next batch inside the current step
    def _step_pipelined(self, audio, v_in, sample_info):
        if self._pgraph is None:
            self._capture_pipelined(audio, v_in, sample_info)           # stages this batch; nothing to report yet
            return None
        for dst, src in zip(self._pstatic[:3], (audio, v_in, sample_info)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._refresh_hyper()
        self._pgraph.replay()
        self.scalars = self._pstatic[3].clone()
        self.losses = self.scalars[:3]
        return self.losses

    def flush_pipeline(self):
        """Runs the model step of the batch staged by the last step() call (its front end is already done); returns its losses."""
        assert self._pgraph is not None
        return self._step_pipelined(*self._pstatic[:3])

    def _capture_pipelined(self, audio, v_in, sample_info):
        nxt = (audio.clone(), v_in.clone(), sample_info.clone())           # static inputs: the batch whose front end runs in the replay
        backup = (self.flat_params.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone())
        bn_state = {k: v.clone() for k, v in self.model.state_dict().items() if 'running' in k}
        main = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for _ in range(2):                                              # warm-up (lazy init, allocator pools, constants upload)
                self._device_step(*nxt, with_optimizer=True)
            x_cur = self._front_end(nxt[0]).clone()                        # the batch the first replay trains on
            x_next = torch.empty_like(x_cur)
        main.wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.flat_params.copy_(backup[0]); self.exp_avg.copy_(backup[1]); self.exp_avg_sq.copy_(backup[2])
        self.model.load_state_dict(bn_state, strict=False)
        cur = (x_cur, nxt[1].clone(), nxt[2].clone())
        fe_stream = torch.cuda.Stream(device=self.device)
        flow = self.model.ae_model.flow_transform

        def launch_next_front_end():
            here = torch.cuda.current_stream(self.device)
            fe_stream.wait_stream(here)
            with torch.cuda.stream(fe_stream):
                self._front_end(nxt[0], out=x_next.view(-1, x_next.shape[-2], x_next.shape[-1]))

        graph = torch.cuda.CUDAGraph()
        before = ops.launches
        flow.on_backward_start = launch_next_front_end
        try:
            with torch.cuda.graph(graph):
                try:
                    scalars = self._model_step(*cur, with_optimizer=True)
                finally:
                    self._drop_prepared()
                torch.cuda.current_stream(self.device).wait_stream(fe_stream)
                cur[0].copy_(x_next); cur[1].copy_(nxt[1]); cur[2].copy_(nxt[2])      # the staged batch becomes the current one
        finally:
            flow.on_backward_start = None
        self.launches_per_step = ops.launches - before
        self.flat_params.copy_(backup[0]); self.exp_avg.copy_(backup[1]); self.exp_avg_sq.copy_(backup[2])
        self.model.load_state_dict(bn_state, strict=False)
        self._pgraph, self._pstatic = graph, (*nxt, scalars, cur)

    # ------------------------------------------------------------------ several ranks: exchange hidden behind compute
    def _step_overlapped(self, audio, v_in, sample_info):
        """Data-parallel step with the whole gradient exchange off the critical path.  Two captured graphs per step:
            A = the mel front end of this step's audio (reads no parameter),   B = forward / losses / backward / packing.
        When the encoder's backward has passed its FC layer, every gradient outside the encoder's convolution stack is final (both
        FC weights, decoder, both flows, regression head: 94 % of the 241 MB).  Graph B packs those there and signals an external
        event; a communication stream reduce-scatters them, applies Adam to this rank's 1/world share, all-gathers the updated
        parameters and refreshes the operand copies of those weights (_exchange_and_update) - under the encoder's convolution
        backward.  After B it does the same for the encoder's stack (15 MB).  The NEXT step launches its graph A before it waits
        for that tail, so what is left of the exchange + update runs under the next front end (0.8 ms of tensor-core work that
        touches no parameter).  Measured at 2 ranks (tools/gpu_timeline_ddp.py): the step is 0.3 ms longer than on one GPU."""
        main = torch.cuda.current_stream(self.device)
        if self._graph is None:
            self._group_counts.copy_(self.controls_criterion.useful_counts(v_in))      # sane values for the warm-up steps of the capture
            self._capture_two(audio, v_in, sample_info)
        st_audio, st_v, st_info = self._static[:3]
        for dst, src in zip((st_audio, st_v, st_info), (audio, v_in, sample_info)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._mark('step start')
        # the categorical normaliser of the gathered batch (54 counts, targets only): exchanged on a side stream under the front end
        aux = self._aux_stream
        aux.wait_stream(main)
        with torch.cuda.stream(aux):
            counts = parallel.global_useful_counts(self.controls_criterion.useful_counts(st_v), self.pg)
            counts.div_(self.world)
            counts_ready = torch.cuda.Event()
            counts_ready.record(aux)
        self._graph_a.replay()                                   # front end -> static spectrogram batch
        self._mark('front end done')
        if self._update_done is not None:
            main.wait_event(self._update_done)                   # previous step's parameters are final from here on
        self._refresh_hyper()
        main.wait_event(counts_ready)
        self._group_counts.copy_(counts)
        counts.record_stream(main)
        self._mark('graph B start (previous update waited for)')
        self._graph.replay()
        scalars = self._static[3].clone()
        end_b = torch.cuda.Event()
        end_b.record(main)
        self._mark('graph B end')
        comm = self._comm_stream
        dist = torch.distributed
        with torch.cuda.stream(comm):
            for part, ready in enumerate((self._dec_ready, self._fc_ready, end_b)):
                comm.wait_event(ready)
                self._mark('comm: phase %d gradients ready' % part)
                self._exchange_and_update(part)
                self._mark('comm: phase %d exchange + Adam + operand refresh done' % part)
            self._update_done = torch.cuda.Event()
            self._update_done.record(comm)
        self.scalars = scalars
        self.losses = scalars[:3]
        return self.losses

    def _shards(self, part):
        """[(lo, hi, my_lo, my_hi)] for the segments of phase `part`: this rank's equal share of each segment."""
        return [(lo, hi) + parallel.shard_of_segment(lo, hi, self.rank, self.world) for lo, hi in self._phase_segments[part]]

    def _exchange_and_update(self, part):
        """Current stream = the communication stream.  Reduce-scatter of the phase's gradient segments (each rank receives the sum of
        its 1/world share), Adam on that share only (1/world of the optimizer's 28 bytes per parameter of HBM traffic, which would
        otherwise compete with the backward pass running beside it), all-gather of the updated parameters, operand copies."""
        if not _DEBUG_SKIP_AR:
            parallel.reduce_scatter_segments_(self.flat_grads, self._phase_segments[part], self.pg)
        self._update_graphs[part][0].replay()                    # Adam on the shards
        if not _DEBUG_SKIP_AR:
            parallel.all_gather_segments_(self.flat_params, self._phase_segments[part], self.pg)
        self._update_graphs[part][1].replay()                    # operand copies of the weights that just changed

    def gather_sharded(self, flat):
        """Completes a flat buffer of which every rank holds its shards only (the reduced gradients, the Adam moments) on all ranks."""
        if getattr(self, '_sharded', False):
            self.finish_updates()
            for seg in self._phase_segments:
                parallel.all_gather_segments_(flat, seg, self.pg)
        return flat

    def _mark(self, name):
        """tools/gpu_timeline_ddp.py: timing events on the current stream while `self._trace` is a list."""
        if getattr(self, '_trace', None) is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.device))
            self._trace.append((name, ev))

    def finish_updates(self):
        """Makes the current stream wait for the parameter update of the last overlapped step (before reading parameters / state)."""
        if getattr(self, '_update_done', None) is not None:
            torch.cuda.current_stream(self.device).wait_event(self._update_done)

    def _capture_two(self, audio, v_in, sample_info):
        static_in = (audio.clone(), v_in.clone(), sample_info.clone())
        backup = (self.flat_params.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone())
        bn_state = {k: v.clone() for k, v in self.model.state_dict().items() if 'running' in k}
        self._make_operands_persistent()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):                                  # warm-up: lazy init, allocator pools, constants upload (no collective)
                self._device_step(*static_in, with_optimizer=False)
            x_static = self._front_end(static_in[0]).clone()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.flat_params.copy_(backup[0]); self.exp_avg.copy_(backup[1]); self.exp_avg_sq.copy_(backup[2])
        self.model.load_state_dict(bn_state, strict=False)
        before = ops.launches
        graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_a):
            self._front_end(static_in[0], out=x_static.view(-1, x_static.shape[-2], x_static.shape[-1]))
        graph_b = torch.cuda.CUDAGraph()
        enc, flow = self.model.ae_model.encoder, self.model.ae_model.flow_transform

        def phase0():
            self._pack_early(0)
            self._dec_ready.record()

        enc.fc_grads_ready_event, enc.before_fc_grads_ready = self._fc_ready, lambda: self._pack_early(1)
        flow.on_backward_start = phase0
        try:
            with torch.cuda.graph(graph_b):
                try:
                    scalars = self._model_step(x_static, static_in[1], static_in[2], with_optimizer=False)
                finally:
                    self._drop_prepared()
        finally:
            enc.fc_grads_ready_event, enc.before_fc_grads_ready, flow.on_backward_start = None, None, None
        # Three phases.  0: final when the latent flow's backward starts (decoder incl. its FC weight, regression flow);  1: final when
        # `_fc_ready` fires (latent flow, the encoder's FC weight);  2: the encoder's convolution stack and its few small vectors (6 %).
        slot_index = {(int(self._offs[i]), self._sizes[i]): i for i in self._direct}
        enc_fc = {slot_index[sl] for sl in self._fc_slots if self._in_encoder[slot_index[sl]]}
        packed = self._early_packed + [[]] * (2 - len(self._early_packed))
        phases = [set(packed[0]) | {i for i in self._direct if not self._in_encoder[i]}, set(packed[1]) | enc_fc]
        phases.append(set(range(len(self.params))) - phases[0] - phases[1])
        total = self.flat_grads.numel()
        self._phase_segments = [parallel.merged_slot_ranges(self._offs, total, ph) for ph in phases]
        self.phase_fractions = [sum(hi - lo for lo, hi in seg) / float(total) for seg in self._phase_segments]
        self.launches_per_step = ops.launches - before + sum(len(seg) for seg in self._phase_segments)      # + the Adam launches on the communication stream
        self.flat_params.copy_(backup[0]); self.exp_avg.copy_(backup[1]); self.exp_avg_sq.copy_(backup[2])
        self.model.load_state_dict(bn_state, strict=False)
        # the update kernels behind each reduction (Adam per segment + the refresh of the operand copies: ~25 small launches) are
        # captured too, so that a step costs the host a handful of calls - with 8 ranks per host the Python threads are the scarce resource
        self._update_graphs = []
        for part in range(3):
            pair = []
            for what in ('adam', 'refresh'):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    if what == 'adam':
                        for lo, hi, mlo, mhi in self._shards(part):
                            self._adam(mlo, mhi)
                    else:
                        self._refresh_operands(conv=('dec', False, 'enc')[part], fc=('dec', 'enc', False)[part], persistent=True)
                pair.append(g)
            self._update_graphs.append(pair)
        self._sharded = True
        # (capturing executed nothing, but be explicit about the state the first real step starts from)
        self.flat_params.copy_(backup[0]); self.exp_avg.copy_(backup[1]); self.exp_avg_sq.copy_(backup[2])
        self._refresh_operands(conv=True, fc=True, persistent=True)
        self._graph_a, self._graph, self._static = graph_a, graph_b, (*static_in, scalars)
        self._update_done = None

    # ------------------------------------------------------------------ host-fed steps with input prefetch
    def prefetch(self, audio_host, v_in_host, sample_info_host):
        """Starts the host -> device copy of the NEXT step's inputs (pinned CPU tensors) on a copy stream, so that it overlaps
        the step currently running on the compute stream.  `step_prefetched()` consumes them."""
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._staged, self._staged_ready, self._staged_free = None, None, None
        if self._staged is None or self._staged[0].shape != audio_host.shape:
            self._staged = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in (audio_host, v_in_host, sample_info_host))
            self._copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        if self._staged_free is not None:           # the staging buffers were last read by the copy into the step's static inputs
            self._copy_stream.wait_event(self._staged_free)
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(self._staged, (audio_host, v_in_host, sample_info_host)):
                dst.copy_(src, non_blocking=True)
            self._staged_ready = torch.cuda.Event()
            self._staged_ready.record(self._copy_stream)

    def step_prefetched(self):
        """One training step on the inputs handed to the last `prefetch()` call."""
        assert getattr(self, '_staged_ready', None) is not None, "call prefetch() first"
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self._staged_ready)
        self._staged_ready = None
        self._staged_free = torch.cuda.Event()
        static = self._pstatic if (self.pipeline_frontend and self._pgraph is not None) else (self._static if self._graph is not None else None)
        if self.use_graph and static is not None:
            for dst, src in zip(static[:3], self._staged):            # device to device, then the staging buffers are free again
                dst.copy_(src, non_blocking=True)
            self._staged_free.record(main)
            return self.step(*static[:3])
        out = self.step(*self._staged)
        self._staged_free.record(main)
        return out

    def losses_to_host_async(self, losses=None):
        """Starts the device -> host copy of a step's loss triple (default: the last step's) into pinned memory on the compute
        stream and returns a handle; `handle.get()` blocks only until THAT copy has finished, so a training loop can log the
        losses of step i while step i+1 is already running instead of draining the GPU every step."""
        scalars = self.scalars if losses is None else losses
        if getattr(self, '_loss_ring', None) is None:
            self._loss_ring = [torch.zeros(len(SCALAR_NAMES), dtype=torch.float32).pin_memory() for _ in range(4)]
            self._loss_ring_pos = 0
        buf = self._loss_ring[self._loss_ring_pos % len(self._loss_ring)]
        self._loss_ring_pos += 1
        buf.zero_()
        buf[:scalars.numel()].copy_(scalars, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return _HostLosses(buf, ev, self.step_count)

    def _capture(self, audio, v_in, sample_info, with_optimizer):
        static_in = (audio.clone(), v_in.clone(), sample_info.clone())
        backup = (self.flat_params.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone())
        bn_state = {k: v.clone() for k, v in self.model.state_dict().items() if 'running' in k}
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):                                  # warm-up: lazy init, allocator pools, constants upload
                self._device_step(*static_in, with_optimizer=with_optimizer)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        # warm-up steps must not count as training: restore parameters, optimizer state and BN statistics
        self.flat_params.copy_(backup[0]); self.exp_avg.copy_(backup[1]); self.exp_avg_sq.copy_(backup[2])
        self.model.load_state_dict(bn_state, strict=False)
        graph = torch.cuda.CUDAGraph()
        before = ops.launches
        self.model.ae_model.encoder.fc_grads_ready_event = self._fc_ready        # recorded (as an external node) by Encoder's backward
        try:
            with torch.cuda.graph(graph):
                losses = self._device_step(*static_in, with_optimizer=with_optimizer)
        finally:
            self.model.ae_model.encoder.fc_grads_ready_event = None
        self.launches_per_step = ops.launches - before
        self.flat_params.copy_(backup[0]); self.exp_avg.copy_(backup[1]); self.exp_avg_sq.copy_(backup[2])
        self.model.load_state_dict(bn_state, strict=False)
        self._graph, self._static = graph, (*static_in, losses)

    @torch.no_grad()
    def infer(self, audio, sample_info=None, full_presets=False):
        """Batched inference audio -> latent -> preset parameters (eval.py:161-182, BASELINE config 5): front end, encoder (+ MIDI
        pitch / velocity concatenation when the model has it, VAE.py:155-165), latent flow and regression flow in eval mode; the decoder
        is not needed for the parameters and is skipped.  full_presets: also run the learnable -> full VST preset conversion
        (data/preset.py:350-369) on the device and return [B, 155]."""
        was_training = self.model.training
        self.finish_updates()
        self.model.eval()
        try:
            B, C, L = audio.shape
            x_in = self.frontend.compute(audio.view(B * C, L), normalize=(self.spec_stats['min'], self.spec_stats['max']))
            x_in = x_in.view(B, C, x_in.shape[-2], x_in.shape[-1])
            ae = self.model.ae_model
            decoder, ae.decoder = ae.decoder, _SkipDecoder()
            try:
                _, _, zk, _, _ = ae(x_in, sample_info)
            finally:
                ae.decoder = decoder
            v_out = self.model.reg_model(zk)
            if full_presets:
                from .data.preset import learnable_to_full_presets
                defaults = getattr(self.idx_helper, 'params_default_values', None) or {}
                return learnable_to_full_presets(self.idx_helper, v_out, defaults)
            return v_out
        finally:
            self.model.train(was_training)

    # ------------------------------------------------------------------ checkpointing (logs/logger.py:199-202)
    def state_dict(self):
        """{'ae_model_state_dict', 'optimizer_state_dict'} in the spirit of the reference's checkpoint file: the optimizer part holds the
        flat Adam moments (one fp32 vector each, in `model.parameters()` order with 16-byte aligned slots) and the step count."""
        self.finish_updates()
        return {'ae_model_state_dict': self.model.state_dict(),
                'optimizer_state_dict': {'step': self.step_count, 'lr': self.lr, 'exp_avg': self.gather_sharded(self.exp_avg).clone(),
                                         'exp_avg_sq': self.gather_sharded(self.exp_avg_sq).clone(),
                                         'slot_offsets': [int(o) for o in self._offs], 'slot_sizes': list(self._sizes)}}

    def load_state_dict(self, state):
        self.finish_updates()
        self.model.load_state_dict(state['ae_model_state_dict'])
        if self._persistent_operands:
            self._refresh_operands(conv=True, fc=True, persistent=True)
        opt = state.get('optimizer_state_dict')
        if opt is not None:
            if list(opt['slot_sizes']) != list(self._sizes):
                raise ValueError("optimizer state was saved for a different parameter layout")
            self.exp_avg.copy_(opt['exp_avg'])
            self.exp_avg_sq.copy_(opt['exp_avg_sq'])
            self.step_count, self.lr = int(opt['step']), float(opt['lr'])


class _SkipDecoder(torch.nn.Module):
    """Stand-in for the decoder during parameter inference: FlowVAE.forward runs unchanged, without the decoder's work."""

    def forward(self, z, dropout_mask=None):
        return None
