#!/usr/bin/env python
"""Benchmark of the preset-gen-vae hot path on B200 (driver contract: see the task statement / DESIGN.md §6).

    python bench.py --gpus 1 --steps K --warmup W                       # this repo's CUDA path, one JSON line
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      # N ranks, NCCL, weak scaling
    python bench.py --impl reference ...                                # the reference algorithm on the host CPU
    ... --strong-scaling                                                # fixed global batch split over the ranks (default: weak scaling)

Workloads (BASELINE.json `configs`):
    train      (default) configs[2]: full ExtendedAE training step from audio - mel front end, conv VAE, latent flow,
               regression flow, losses, backward, Adam - C=1, per-GPU batch 160 (config.py:80); metric = samples/s
    train_c6   configs[3]: the same step on multi-note input, 6 MIDI notes stacked as spectrogram channels (config.py:35-37)
    frontend   configs[1]: STFT + mel front end alone, 256 clips
    inference  configs[4]: audio -> latent -> preset parameters, batch 1024
`value` is timed with inputs resident in HBM; `e2e` goes through the public API with pinned HOST buffers (H2D of
the step's audio / targets and D2H of the losses inside the timed region).  `roofline` describes the kernel that dominates the
live run and reports the larger of its two floors (algorithmic flops / tensor peak, algorithmic bytes / HBM peak) as the bound.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = {'train': 'train samples/sec', 'train_c6': 'train samples/sec', 'frontend': 'front-end clips/sec', 'inference': 'inference samples/sec'}
SIX_NOTES = ((40, 85), (50, 85), (60, 42), (60, 85), (60, 127), (70, 85))          # config.py:35 of the reference
FRONTEND_FLOP_PER_CLIP = 820.63e6          # dense DFT (729.13 MFLOP) + dense mel (91.50 MFLOP)
DFT_FLOP_PER_CLIP, MEL_FLOP_PER_CLIP = 729.13e6, 91.50e6


# C entry point -> CUDA kernel it launches (one kernel serves several entry points)
KERNEL_OF = {'pgv_conv_cl_fwd': 'conv_cl_kernel', 'pgv_conv_cl_dgrad': 'conv_cl_kernel', 'pgv_conv_cl_wgrad': 'conv_cl_kernel',
             'pgv_conv_cl_fwd_bn': 'conv_cl_kernel', 'pgv_conv_cl_dgrad_bn': 'conv_cl_kernel',
             'pgv_conv2d_fwd_tf32': 'conv_tc_kernel', 'pgv_conv2d_dgrad_tf32': 'conv_tc_kernel', 'pgv_conv2d_wgrad_tf32': 'conv_tc_kernel',
             'pgv_linear_fwd_tf32': 'conv_tc_kernel', 'pgv_linear_dgrad_tf32': 'conv_tc_kernel', 'pgv_linear_wgrad_tf32': 'conv_tc_kernel',
             'pgv_linear_cl_fwd': 'conv_cl_kernel', 'pgv_linear_cl_dgrad': 'conv_cl_kernel', 'pgv_linear_cl_wgrad': 'conv_cl_kernel',
             'pgv_linear_cs_fwd': 'colslice_gemm_kernel', 'pgv_linear_cs_dgrad': 'colslice_gemm_kernel', 'pgv_linear_bn_fwd': 'colslice_gemm_kernel',
             'pgv_linear_dgrad_bn_bwd': 'colslice_gemm_kernel',
             'pgv_gemm_f32': 'gemm_f32_small_kernel', 'pgv_linear_wgrad_f32': 'gemm_f32_small_kernel',
             'pgv_flow_program': 'flow_program_kernel'}


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of all launches of `kernel` in one training step, from the committed ncu
    launch list of tools/run_train_once.py (profiles/traffic_r02.json, written by tools/summarize_launches.py); None if absent."""
    path = os.path.join(ROOT, 'profiles', 'traffic_r02.json')
    if not os.path.exists(path):
        return None
    rec = json.load(open(path)).get(kernel)
    return None if rec is None else int(rec['dram_bytes'])


def ncu_kernel_times():
    """{kernel: summed gpu__time_duration (us) of its launches in one training step} from profiles/traffic_r02.json, or {}."""
    path = os.path.join(ROOT, 'profiles', 'traffic_r02.json')
    if not os.path.exists(path):
        return {}
    return {k: float(v.get('us', 0.0)) for k, v in json.load(open(path)).items()}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p['hbm_gbs'], bf16_burst=p['bf16_tflops'], bf16_sustained=p['bf16_tflops_sustained'], src='measured')
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.rows = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (('hw_slowdown', 5), ('hw_thermal_slowdown', 6), ('sw_thermal_slowdown', 7), ('sw_power_cap', 8)):
                if len(r) > col and r[col].lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def dist_env():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def default_batch(workload):
    return {'train': 160, 'train_c6': 160, 'frontend': 256, 'inference': 1024}[workload]


# ---------------------------------------------------------------------------------------------------- reference arm
def oracle_step_factory(workload, batch):
    """The reference algorithm on the host CPU (oracle port: nflows / librosa restated, see oracle/__init__.py).
    Returns (callable running ONE step on `batch` samples, description)."""
    from oracle import frontend as ofe, losses as oloss, model as omodel
    from preset_gen_vae_b200 import synthetic
    helper, m_cfg, t_cfg = model_configs(workload, batch)
    C = m_cfg.input_tensor_size[1]
    audio = synthetic.make_audio(min(batch, 64), C, seed=0)
    audio = audio.repeat((batch + audio.shape[0] - 1) // audio.shape[0], 1, 1)[:batch].contiguous()
    st = synthetic.SPEC_STATS
    if workload == 'frontend':
        def step():      # one clip at a time, as the reference DataLoader worker does (abstractbasedataset.py:124-134)
            return ofe.batch_front_end(audio, 1024, 256, -120.0, 257, st['min'], st['max'])
        return step, 'oracle port of utils/audio.py MelSpectrogram, per-clip loop'
    torch.manual_seed(0)
    ext = omodel.build_extended_ae_model(m_cfg, t_cfg, helper)[3]
    v_in = synthetic.make_preset_targets(helper, batch, seed=0)
    info = synthetic.make_sample_info(batch)
    if workload == 'inference':
        ext.eval()

        def step():
            with torch.no_grad():
                x = ofe.batch_front_end(audio, 1024, 256, -120.0, 257, st['min'], st['max'])
                ml = ext.ae_model.encoder(x)
                zk, _ = ext.ae_model.flow_transform(ml[:, 0, :])
                return ext.reg_model(zk)
        return step, 'oracle port: front end + encoder + flows + regression head (eval)'
    ext.train()
    opt = torch.optim.Adam(ext.parameters(), lr=t_cfg.initial_learning_rate, weight_decay=t_cfg.weight_decay, betas=t_cfg.adam_betas)

    def step():          # train.py:204-248 without DataParallel
        x = ofe.batch_front_end(audio, 1024, 256, -120.0, 257, st['min'], st['max'])
        opt.zero_grad()
        _, losses, total = oloss.train_step_losses(ext, x, v_in, info, None, beta=t_cfg.beta)
        total.backward()
        opt.step()
        return total
    return step, 'oracle port of train.py:204-248 (front end per clip + fwd + losses + bwd + torch Adam)'


def run_cpu(workload, batch, budget_s, min_steps=2, max_steps=50):
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    step, what = oracle_step_factory(workload, batch)
    step()                                            # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < min_steps or (time.perf_counter() < t_end and len(times) < max_steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return dict(value=batch / med, unit='samples/s', cores=threads, kind='port',
                sample='%d steps of batch %d (%s), median %.3f s/step' % (len(times), batch, what, med)), med


def main_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    batch = args.batch_per_gpu or default_batch(args.workload)
    if args.strong_scaling:
        batch = max(4, batch // max(args.gpus, 1) // 4 * 4)
    cpu_batch = args.cpu_batch or batch                # the same per-GPU batch as our arm: same workload string, same config object
    budget = max(10.0, min(150.0, 12.0 * max(args.steps, 1)))
    cb, med = run_cpu(args.workload, cpu_batch, budget)
    line = {'impl': 'reference', 'metric': METRIC[args.workload], 'value': cb['value'], 'unit': 'samples/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': med * 1e3, 'higher_is_better': True,
            'scaling': 'strong' if args.strong_scaling else 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': line_config(args.workload, cpu_batch, args.gpus, not args.no_graph),
            'note': 'reference algorithm (oracle port) on the host CPU of rank 0: one replica of the per-GPU batch, no GPU',
            'cpu_baseline': cb, 'e2e': {'value': cb['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    emit(line)


def model_configs(workload, batch):
    from preset_gen_vae_b200 import config as pcfg
    from preset_gen_vae_b200.data.preset import DexedLearnableLayout
    helper = DexedLearnableLayout().preset_indexes_helper
    if workload == 'train_c6':
        m_cfg, t_cfg = pcfg.make_default(minibatch_size=batch, midi_notes=SIX_NOTES, stack_spectrograms=True)
    else:
        m_cfg, t_cfg = pcfg.make_default(minibatch_size=batch)
    pcfg.apply_dataset_dims(m_cfg, helper)
    return helper, m_cfg, t_cfg


def line_config(workload, batch, gpus, cuda_graph):
    """The `config` object of the JSON line: identical for both arms (the driver compares them)."""
    return {'workload': workload_name(workload, batch), 'global_batch': gpus * batch, 'parallelism': 'dp%d' % gpus,
            'step': 'one timed step = one front end + one forward/backward/Adam over a batch; on one GPU the front end of batch i+1 runs '
                    'inside the step of batch i (software pipeline), on several GPUs between the steps under the gradient exchange',
            'l2': 'per-step working set (activations + 241 MB of parameters, > 1 GB) exceeds the 126 MB L2; no flush needed',
            'cuda_graph': bool(cuda_graph)}


def workload_name(workload, batch):
    return {'train': 'full ExtendedAE train step from audio (mel front end + speccnn8l1_bn conv VAE + realnvp_6l300 latent flow + '
                     'flow_realnvp_6l300 regression + losses, fwd+bwd+Adam), C=1, dim_z=610, per-GPU batch %d' % batch,
            'train_c6': 'full ExtendedAE train step from audio, 6 MIDI notes stacked as spectrogram channels (C=6: shared per-channel CNNs, '
                        '4x4 mixer 1536->768, 67.1 M parameters), dim_z=610, per-GPU batch %d' % batch,
            'frontend': 'STFT + mel spectrogram front end alone, %d synthetic clips of 88576 samples (n_fft 1024, hop 256, 257 mel)' % batch,
            'inference': 'batched inference audio -> latent -> Dexed preset parameters, batch %d per GPU' % batch}[workload]


# ---------------------------------------------------------------------------------------------------- our arm
def main_ours(args):
    rank, local_rank, world = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    pg = None
    if world > 1:
        # (NCCL's debug output, if the environment asks for it, stays where the environment points it; stdout carries the JSON line only,
        # see _guard_stdout)
        torch.distributed.init_process_group('nccl', device_id=dev)
        pg = torch.distributed.group.WORLD
    from preset_gen_vae_b200 import _lib, config as pcfg, synthetic
    from preset_gen_vae_b200.data.preset import DexedLearnableLayout
    from preset_gen_vae_b200.model import ops
    from preset_gen_vae_b200.train import TrainStep
    from preset_gen_vae_b200.utils.audio import MelSpectrogram
    B = args.batch_per_gpu or default_batch(args.workload)
    if args.strong_scaling:                          # the fixed global batch of the workload split over the ranks
        B = max(4, B // world // 4 * 4)
    helper, m_cfg, t_cfg = model_configs(args.workload, B)
    C = m_cfg.input_tensor_size[1]
    audio_h = synthetic.make_audio(min(B, 64), C, seed=rank)                 # CPU synthesis is slow: tile 64 distinct clips
    audio_h = audio_h.repeat((B + audio_h.shape[0] - 1) // audio_h.shape[0], 1, 1)[:B].contiguous().pin_memory()
    v_in_h = synthetic.make_preset_targets(helper, B, seed=rank).pin_memory()
    info_h = synthetic.make_sample_info(B).pin_memory()
    audio, v_in, info = audio_h.to(dev), v_in_h.to(dev), info_h.to(dev)
    st = synthetic.SPEC_STATS

    trainer = None
    if args.workload == 'frontend':
        mel = MelSpectrogram(1024, 256, -120.0, 257, 22050, device=dev)
        a2 = audio.view(B, -1)
        launches = _lib.lib().pgv_frontend_launch_count(257)

        def dev_step():
            return mel.compute(a2, normalize=(st['min'], st['max']))

        def e2e_step():
            return mel.compute_host(audio_h.view(B, -1), normalize=(st['min'], st['max']), device=dev)
        h2d, d2h = audio_h.numel() * 4, B * 257 * 347 * 4
    else:
        trainer = TrainStep(m_cfg, t_cfg, helper, device=dev, process_group=pg, use_cuda_graph=not args.no_graph,
                            pipeline_frontend=not args.no_pipeline)
        if args.workload in ('train', 'train_c6'):
            def dev_step():
                return trainer.step(audio, v_in, info)

            trainer.prefetch(audio_h, v_in_h, info_h)

            pending = []

            def e2e_step():
                # public host-fed API: the step runs on the batch staged by the previous call while the pinned-host -> device copy of
                # the next batch (one full copy per step, inside the timed region) overlaps it on the copy stream.  Every step's
                # losses are read by the host: copied to pinned memory behind the step and fetched while the NEXT step runs
                # (the usual non-blocking loss logging); the last ones are fetched by e2e_finish() inside the timed region.
                trainer.step_prefetched()
                trainer.prefetch(audio_h, v_in_h, info_h)
                if trainer.scalars is not None:
                    pending.append(trainer.losses_to_host_async())
                return pending.pop(0).get() if len(pending) > 1 else None

            def e2e_finish():
                while pending:
                    last = pending.pop(0).get()
                assert bool(torch.isfinite(last).all())
            h2d, d2h = (audio_h.numel() + v_in_h.numel() + info_h.numel()) * 4, 12
        else:
            def dev_step():
                return trainer.infer(audio)

            # host-fed inference: the pinned-host -> device copy of the NEXT batch (one full copy per step, inside the timed region)
            # runs on a copy stream under the batch being processed; the preset parameters are read back every step
            copy_stream = torch.cuda.Stream(device=dev)
            staged, staged_ready, turn = [torch.empty_like(audio) for _ in range(2)], [None, None], [0]

            def stage(i):
                copy_stream.wait_stream(torch.cuda.current_stream(dev))       # buffer i was last read by an earlier infer()
                with torch.cuda.stream(copy_stream):
                    staged[i].copy_(audio_h, non_blocking=True)
                    staged_ready[i] = torch.cuda.Event()
                    staged_ready[i].record(copy_stream)

            stage(0)

            def e2e_step():
                i = turn[0] & 1
                turn[0] += 1
                torch.cuda.current_stream(dev).wait_event(staged_ready[i])
                stage(i ^ 1)
                return trainer.infer(staged[i]).cpu()
            h2d, d2h = audio_h.numel() * 4, B * 610 * 4
        launches = None

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        dev_step()
    before = ops.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = timed(dev_step, args.steps)
    if total_ms < 400.0:
        # the timed region is too short for nvidia-smi's sampling period: keep the SAME step running (untimed) until the sampler has
        # seen ~0.5 s of it, so that the clock / throttle record describes this load; every rank does the same number of steps
        extra = int(min(5000, max(8, 500.0 / max(total_ms / args.steps, 1e-3))))
        for _ in range(extra):
            dev_step()
        torch.cuda.synchronize(dev)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None and total_ms < 400.0:
        clocks['note'] = 'timed region %.0f ms: sampling continued over %d more untimed steps of the same workload' % (total_ms, extra)
    if launches is None:
        launches = getattr(trainer, 'launches_per_step', None) or (ops.launches - before) // max(args.steps, 1)
    ms_per_step = total_ms / args.steps
    value = world * B / ms_per_step * 1e3
    e2e_fin = locals().get('e2e_finish')
    for _ in range(2):
        e2e_step()
    if e2e_fin is not None:
        e2e_fin()
    e2e_ms = timed(e2e_step, args.steps, e2e_fin) / args.steps
    e2e = {'value': world * B / e2e_ms * 1e3, 'unit': 'samples/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
           'ms_per_step': e2e_ms}
    if args.workload in ('train', 'train_c6'):
        e2e['note'] = ('TrainStep.prefetch / step_prefetched: every timed step issues one pinned-host -> device copy of a full batch (the next '
                       "step's inputs, on a copy stream, overlapping the running step) and reads every step's three losses on the host (async copy to pinned "
                       "memory behind the step, fetched while the next step runs; the last step's are fetched before the region ends)")
    elif args.workload == 'inference':
        e2e['note'] = ('TrainStep.infer on batches from pinned host memory: every timed step issues one full host -> device copy (the next '
                       'batch, on a copy stream, under the batch being processed) and reads the predicted preset parameters back to the host')

    # ---- live roofline of the dominant kernel family: eager steps with CUDA events around every C entry point ----
    pk = peaks()
    roofline, breakdown = None, None
    if args.workload == 'frontend':
        if rank == 0:
            # the three front-end kernels ARE the timed step: CUDA events around the K back-to-back steps of the timed region
            ms = ms_per_step
            ach = FRONTEND_FLOP_PER_CLIP * B / (ms * 1e-3) / 1e12
            roofline = {'kernel': 'gemm_tf32_kernel<DftProblem> + <MelProblem> (whole front end)', 'bound': 'tensor', 'achieved': ach,
                        'peak': pk['bf16_burst'] / 2, 'unit': 'TFLOP/s', 'frac': ach / (pk['bf16_burst'] / 2), 'traffic': None,
                        'note': 'algorithmic dense flops 820.63 MFLOP/clip; the 3xTF32 split executes 3x that on the tensor pipe; '
                                'peak = TF32 dense = half of the %s bf16 burst figure' % pk['src']}
    else:
        # EVERY rank runs the instrumented eager steps (they contain the gradient all-reduce); rank 0 reports.
        # The instrumented pass is eager and single-stream (no decoder side stream), so that the events around each entry point
        # measure that kernel alone and not the time it shared the GPU with a concurrent branch.
        trainer_graph, trainer_side, trainer_pipe = trainer.use_graph, trainer._side, trainer.pipeline_frontend
        trainer.use_graph, trainer._side, trainer.pipeline_frontend = False, None, False
        ops.use_wgrad_fork = False                      # same reason: weight gradients on the measuring stream, not on their child stream
        dev_step()
        torch.cuda.synchronize(dev)
        # keep the GPU busy while the CPU enqueues the instrumented steps, so that the CUDA events around each entry
        # point measure device execution only (without this, short kernels are dominated by CPU launch gaps)
        blocker = torch.empty(16384, 16384, device=dev)
        for _ in range(6):
            torch.mm(blocker, blocker)
        ops.start_profile()
        for _ in range(2):
            dev_step()
        prof = ops.stop_profile()
        del blocker
        trainer.use_graph, trainer._side, trainer.pipeline_frontend = trainer_graph, trainer_side, trainer_pipe
        ops.use_wgrad_fork = True
        if rank == 0:
            tot = sum(v['ms'] for v in prof.values())
            breakdown = {k: round(v['ms'] / 2, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])[:10]}
            # entry points -> the CUDA kernel they launch; the dominant KERNEL (all its launches of one step) is the roofline subject
            fam = {}
            for name, v in prof.items():
                f = fam.setdefault(KERNEL_OF.get(name, name), dict(ms=0.0, flops=0, bytes=0, calls=0))
                f['ms'] += v['ms']; f['flops'] += v['flops']; f['bytes'] += v['bytes']; f['calls'] += v['calls']
            # Which kernel dominates the step: the live per-entry-point CUDA-event totals of THIS run (each call bracketed by its own
            # pair of events while the GPU is kept busy, so launch gaps between calls are not counted).
            ranked = sorted(fam.items(), key=lambda kv: -kv[1]['ms'])
            name, top = ranked[0]
            per_ms = top['ms'] / 2
            # (the ncu launch list is one B=160 training step: its DRAM bytes only describe that workload)
            traffic = measured_traffic(name) if args.workload == 'train' else None
            n_launch = max(1, top['calls'] // 2)
            if top['flops'] > 0:
                tensor = name in ('conv_cl_kernel', 'conv_tc_kernel')
                peak_t = pk['bf16_sustained'] / 2
                ach_t = top['flops'] / 2 / (per_ms * 1e-3) / 1e12
                ach_h = top['bytes'] / 2 / (per_ms * 1e-3) / 1e9
                frac_t, frac_h = ach_t / peak_t, ach_h / pk['hbm_gbs']
                # the binding roofline of the kernel's launches taken together is the larger of its two floors (flops / tensor peak,
                # algorithmic bytes / HBM peak): the 8..32-channel layers make conv_cl_kernel HBM-class in aggregate
                hbm_bound = tensor and frac_h > frac_t
                roofline = {'kernel': name, 'bound': 'hbm' if hbm_bound else 'tensor', 'achieved': ach_h if hbm_bound else ach_t,
                            'peak': pk['hbm_gbs'] if hbm_bound else peak_t, 'unit': 'GB/s' if hbm_bound else 'TFLOP/s',
                            'frac': frac_h if hbm_bound else frac_t, 'frac_tensor': frac_t, 'frac_hbm': frac_h,
                            'achieved_tflops': ach_t, 'achieved_algorithmic_gbs': ach_h,
                            'traffic': None if traffic is None else traffic / n_launch, 'traffic_per_step': traffic,
                            'launches_per_step': n_launch, 'ms_per_step': per_ms, 'us_per_launch': per_ms * 1e3 / n_launch,
                            'share_of_step': top['ms'] / tot, 'algorithmic_bytes_per_step': top['bytes'] // 2,
                            'algorithmic_bytes_per_launch': top['bytes'] // 2 // n_launch, 'algorithmic_flops_per_launch': top['flops'] // 2 // n_launch,
                            'note': ('tcgen05 kind::tf32 kernel' if tensor else 'CUDA-core fp32 kernel (not on the tensor pipe)') +
                                    '; achieved = algorithmic bytes (or flops) of all its launches in one step / their summed CUDA-event time '
                                    '(events on the launching stream, GPU kept busy) = per-launch average / per-launch average duration; bound = the '
                                    'larger of the two floors; tensor peak = TF32 dense = half of the %s sustained bf16 figure, HBM peak = the %s copy '
                                    'bandwidth; traffic = dram bytes per launch (average over the same launches of one step under ncu, '
                                    'profiles/traffic_r02.json)' % (pk['src'], pk['src'])}
            else:
                ach = top['bytes'] / 2 / (per_ms * 1e-3) / 1e9
                roofline = {'kernel': name, 'bound': 'hbm', 'achieved': ach, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': ach / pk['hbm_gbs'],
                            'traffic': traffic, 'share_of_step': top['ms'] / tot, 'note': 'peak = %s copy bandwidth' % pk['src']}
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:       # reported at N=1 only
        cpu_baseline, _ = run_cpu(args.workload, args.cpu_batch or B, budget_s=args.cpu_budget)
    if rank == 0:
        line = {'metric': METRIC[args.workload], 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
                'scaling': 'strong' if args.strong_scaling else 'weak', 'vs_baseline': None,
                'dtype': 'f32 storage; tf32 tensor-core products where the layer runs on tcgen05, fp32 elsewhere',
                'data': 'synthetic', 'config': line_config(args.workload, B, world, not args.no_graph),
                'e2e': e2e, 'gpu_launches': int(launches * args.steps), 'gpu_launches_per_step': int(launches), 'clocks': clocks,
                'roofline': roofline, 'cpu_baseline': cpu_baseline}
        if breakdown is not None:
            line['ms_by_entry_point_eager'] = breakdown
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


_json_out = None


def _guard_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner to fd 1 when
    NCCL_DEBUG is set in the environment), so fd 1 is pointed at stderr for the whole run and the JSON line goes to a private
    duplicate of the original stdout."""
    global _json_out
    sys.stdout.flush()
    _json_out = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def emit(line):
    out = _json_out if _json_out is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='train', choices=['train', 'train_c6', 'frontend', 'inference'])
    ap.add_argument('--batch-per-gpu', type=int, default=0)
    ap.add_argument('--cpu-batch', type=int, default=0, help='batch of the bounded CPU sample (0 = the per-GPU batch of the workload)')
    ap.add_argument('--cpu-budget', type=float, default=20.0, help='seconds of CPU baseline work')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--strong-scaling', action='store_true', help='split the workload\'s per-GPU batch over the ranks (fixed global batch) instead of weak scaling')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-pipeline', action='store_true', help='one GPU: do not software-pipeline the front end (TrainStep(pipeline_frontend=False))')
    args = ap.parse_args()
    if args.impl == 'reference':
        main_reference(args)
    else:
        main_ours(args)


if __name__ == '__main__':
    main()
