#!/usr/bin/env python
"""bench.py -- StyleGAN2 G+D training throughput (images/sec) on B200, BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched per rank by torch.distributed.run)
    python bench.py --impl reference ...                      (the reference's CPU impl='ref' algorithm, oracle port)

Workload (config.workload): BASELINE.json configs[1] -- "StyleGAN2 config-f 1024^2 G+D forward-backward fp32,
batch 32 on 1 B200": one step = one training iteration of the upstream loop (Gmain + Dmain every iteration,
Greg every 4th, Dreg every 16th, Adam steps, G_ema), batch 32 PER GPU (weak scaling), run as `batch_gpu`-sized
accumulation rounds exactly like training_loop.py:495-502.  Synthetic data, random-init weights.

One JSON line on stdout (rank 0):
  value      images/sec, inputs resident in HBM when the timed region starts (CUDA events, max over ranks)
  e2e        images/sec through the public API with HOST inputs: every step copies that step's uint8 image batch
             and latents from pinned host memory, and reads the step's loss scalars back
  roofline   the dominant kernel family (conv2d fwd/dgrad/wgrad): algorithmic FLOPs / CUDA-event time of every
             conv launch inside the timed region, against the measured TF32 peak (= 1/2 of the measured bf16 peak)
  cpu_baseline  the oracle port of the reference's impl='ref' CPU path, timed on this box's host cores (rank 0, N=1)
"""
import os
import sys
import json
import time
import argparse
import threading
import subprocess

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'ga-gan_b200')
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np   # noqa: E402
import torch         # noqa: E402

METRIC = 'stylegan2_g_d_train_images_per_sec'
UNIT = 'img/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=16)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--res', type=int, default=1024)
    ap.add_argument('--cfg', default='stylegan2')
    ap.add_argument('--batch', type=int, default=32, help='images per GPU per iteration')
    ap.add_argument('--batch-gpu', type=int, default=32, help='images per accumulation round (training_loop.py:495-502); one round of 32 fits '
                    "the B200's 180 GB in fp32 and keeps the low-resolution layers' tiles full (4 -> 22, 8 -> 24, 16 -> 26, 32 -> 28+ img/s)")
    ap.add_argument('--prec', default='auto', choices=['auto', 'simt', 'tf32x1', 'tf32x3'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-res', type=int, default=0, help='resolution of the CPU sample (0 = same as --res)')
    ap.add_argument('--cpu-batch', type=int, default=2)
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d['hbm_gbs'], bf16_burst=d['bf16_tflops'], bf16_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    source='MEASURED_PEAKS.json')
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200',
                                          '-i', str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [v.strip() for v in ln.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


# ----------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's impl='ref' path (the reference is Python and cannot travel
# to the GPU box; oracle/ is pinned against it by tests/golden/make_golden.py).


def cpu_reference_iteration(res, cfg, batch, seed=0):
    """One amortised training iteration on the host cores; returns (images/sec, seconds per phase)."""
    from oracle import networks_ref as NR
    from training.training_loop import CONFIGS
    spec = CONFIGS[cfg]
    cb = int(spec['fmaps'] * 32768)
    torch.set_num_threads(os.cpu_count())
    gen = torch.Generator().manual_seed(seed)
    PG = {k: v.requires_grad_(v.dtype.is_floating_point and 'noise_const' not in k and 'w_avg' not in k)
          for k, v in NR.init_G_params(res, cb, 512, 512, 512, spec['map'], generator=gen).items()}
    PD = {k: v.requires_grad_(True) for k, v in NR.init_D_params(res, cb, 512, generator=gen).items()}
    gl = [v for v in PG.values() if v.requires_grad]
    dl = list(PD.values())
    optG = torch.optim.Adam(gl, lr=spec['lrate'] * 0.8, betas=(0.0, 0.99 ** 0.8), eps=1e-8)
    optD = torch.optim.Adam(dl, lr=spec['lrate'] * 16 / 17, betas=(0.0, 0.99 ** (16 / 17)), eps=1e-8)
    z = torch.randn(batch, 512, generator=gen)
    real = torch.rand(batch, 3, res, res, generator=gen) * 2 - 1
    mb = spec['mbstd']
    t = {}

    def phase(name, fn, params, opt):
        t0 = time.perf_counter()
        for p in params:
            p.grad = None
        loss = fn()
        loss.backward()
        opt.step()
        t[name] = time.perf_counter() - t0

    phase('Gmain', lambda: NR.loss_Gmain(PG, PD, z, res, mb, num_layers=spec['map']), gl, optG)
    phase('Dmain', lambda: NR.loss_Dmain(PG, PD, z, real, res, mb, num_layers=spec['map']), dl, optD)
    phase('Greg', lambda: NR.loss_Gpl(PG, z, res, torch.zeros([]), num_layers=spec['map'])[0] * 4, gl, optG)
    phase('Dreg', lambda: NR.loss_Dr1(PD, real, res, spec['gamma'], mb) * 16, dl, optD)
    t_iter = t['Gmain'] + t['Dmain'] + t['Greg'] / 4 + t['Dreg'] / 16
    return batch / t_iter, t


def run_reference_arm(args, rank):
    """`--impl reference`: the reference's CPU algorithm (oracle port) on all host threads, rank 0 only."""
    if rank != 0:
        return
    res = args.cpu_res or args.res
    budget_s = 240.0
    t_start = time.perf_counter()
    vals, phases, done_warm, done = [], None, 0, 0
    for i in range(args.warmup + args.steps):
        v, ph = cpu_reference_iteration(res, args.cfg, args.cpu_batch, seed=i)
        if i >= args.warmup or (time.perf_counter() - t_start) > budget_s * 0.5:
            vals.append(v); phases = ph; done += 1
        else:
            done_warm += 1
        if (time.perf_counter() - t_start) > budget_s and vals:
            break
    value = float(np.mean(vals))
    sample = (f'{done} timed + {done_warm} warm-up amortised iterations (Gmain+Dmain+Greg/4+Dreg/16 incl. Adam) at batch '
              f'{args.cpu_batch}, {res}x{res} {args.cfg} fp32, oracle port of impl=ref; wall budget {budget_s:.0f}s')
    line = dict(metric=METRIC, value=value, unit=UNIT, impl='reference', n_gpus=args.gpus, steps=done, warmup=done_warm,
                ms_per_step=1000.0 * args.cpu_batch / value, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic',
                # the same workload as the GPU arm (its `config.workload`, `global_batch`); what was actually timed is `cpu_baseline.sample`
                config=dict(workload=f'StyleGAN2 {args.cfg} (config-f) {args.res}x{args.res} G+D train iteration fp32, batch {args.batch}/GPU '
                                     f'in rounds of {min(args.batch_gpu, args.batch)}', global_batch=args.batch * max(args.gpus, 1),
                            parallelism=f'dp{max(args.gpus, 1)}', reference_sample=f'batch {args.cpu_batch} at {res}x{res} on the host cores, scaled per image'),
                cpu_baseline=dict(value=value, unit=UNIT, cores=os.cpu_count(), kind='port', sample=sample,
                                  phase_seconds={k: round(v, 3) for k, v in phases.items()}),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference_arm(args, rank)
        return

    import torch.distributed as dist
    from torch_utils import custom_ops
    from training import training_loop

    assert torch.cuda.is_available(), 'bench.py needs a B200; there is no CPU path (use --impl reference for the CPU arm)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)'
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    custom_ops.verbosity = 'none'
    custom_ops.conv_precision = dict(auto=custom_ops.PREC_AUTO, simt=custom_ops.PREC_FP32_SIMT, tf32x1=custom_ops.PREC_TF32X1,
                                     tf32x3=custom_ops.PREC_TF32X3)[args.prec]

    spec = training_loop.CONFIGS[args.cfg]
    torch.manual_seed(0 * world + rank)                       # training_loop.py:204-205
    G, D = training_loop.build_networks(args.res, args.cfg, device=dev)
    step = training_loop.TrainingStep(G, D, batch_size=args.batch * world, batch_gpu=min(args.batch_gpu, args.batch), device=dev,
                                      lrate=spec['lrate'], r1_gamma=spec['gamma'], ema_kimg=spec['ema'], rank=rank, num_gpus=world)
    n_phases = len(step.phases)

    # synthetic inputs: device-resident for `value`, pinned host uint8 + host latents for `e2e`
    real_dev = torch.rand(args.batch, 3, args.res, args.res, device=dev) * 2 - 1      # training_loop.py:441 range
    real_host = torch.randint(0, 256, (args.batch, 3, args.res, args.res), dtype=torch.uint8).pin_memory()
    z_host = torch.randn(n_phases, args.batch, 512).pin_memory()
    h2d_bytes = real_host.numel() + z_host.numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, host_inputs):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        step.cur_it = 0                       # Greg fires ceil(K/4) times, Dreg ceil(K/16) times: never under-counted
        d2h = 0
        e0.record()
        for _ in range(n_steps):
            if host_inputs:
                real = real_host.to(dev, non_blocking=True).to(torch.float32) / 127.5 - 1
                zs = z_host.to(dev, non_blocking=True)
                out = step.run(real, zs)
                host_vals = torch.stack([v for v in out.values()]).cpu()      # the step's loss scalars -> host (syncs)
                d2h = host_vals.numel() * 4
            else:
                step.run(real_dev)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), d2h

    for _ in range(max(args.warmup, 0)):                      # warm-up covers all four phases (cur_it = 0 fires both regs)
        step.cur_it = 0
        step.run(real_dev)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    custom_ops.conv_profile = []
    launches0 = custom_ops.launch_count()
    ms_total, _ = timed(args.steps, host_inputs=False)
    launches = custom_ops.launch_count() - launches0
    prof, custom_ops.conv_profile = custom_ops.conv_profile, None
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, d2h_bytes = timed(args.steps, host_inputs=True)

    imgs = args.batch * world * args.steps
    value = imgs / (ms_total / 1000.0)
    e2e_value = imgs / (ms_e2e / 1000.0)

    # roofline of the dominant kernel, from the per-launch CUDA events recorded inside the timed region: conv_tc_kernel, the
    # tcgen05 implicit-GEMM convolution that serves every forward and data-gradient contraction (by_kind lists the others)
    peaks = measured_peaks()
    by_kind = {}
    for kind, flops, prec, a, b in prof:
        k = f'{kind}/{ {0: "fp32_ffma", 1: "tf32x1", 3: "tf32x3"}.get(prec, prec) }'
        d = by_kind.setdefault(k, [0, 0.0, 0.0])
        d[0] += 1; d[1] += flops; d[2] += a.elapsed_time(b)
    dom = [v for k, v in by_kind.items() if k in ('conv/tf32x3', 'convT/tf32x3', 'conv/tf32x1', 'convT/tf32x1')]
    dom_launches = sum(v[0] for v in dom); dom_flops = sum(v[1] for v in dom); dom_ms = sum(v[2] for v in dom)
    all_conv_ms = sum(v[2] for v in by_kind.values())
    tf32_peak = peaks['bf16_sustained'] / 2.0
    achieved = dom_flops / (dom_ms / 1000.0) / 1e12 if dom_ms > 0 else 0.0
    # DRAM bytes per launch of the same kernel from the committed ncu --set full capture of this command (profiles/): bench.py cannot
    # run a profiler itself, so the figure is read from the summary that tools/ncu_traffic.py wrote; null if there is none
    traffic, traffic_note = None, 'no ncu capture committed'
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.isfile(tpath):
        try:
            tj = json.load(open(tpath))
            traffic, traffic_note = tj['conv_tc_kernel']['dram_bytes_per_launch'], tj['conv_tc_kernel']['note']
        except Exception as e:      # a malformed summary must not break the bench line
            traffic_note = f'unreadable {tpath}: {e}'
    roofline = dict(bound='tensor', achieved=achieved, peak=tf32_peak, unit='TFLOP/s', frac=achieved / tf32_peak, traffic=traffic,
                    traffic_note=traffic_note,
                    kernel='conv_tc_kernel (tcgen05 implicit-GEMM conv: forward + data gradient of every conv layer)',
                    launches=dom_launches, avg_launch_ms=(dom_ms / dom_launches if dom_launches else None),
                    algorithmic_flops_per_launch=(dom_flops / dom_launches if dom_launches else None),
                    share_of_step=dom_ms / ms_total, all_conv_kernels_share_of_step=all_conv_ms / ms_total,
                    peak_source=f'{peaks["source"]}: bf16 sustained {peaks["bf16_sustained"]} TF/s / 2 (TF32 dense = half of bf16)',
                    note='fp32 parity needs 3 TF32 products per MAC (hi*hi + hi*lo + lo*hi): the tensor pipe does 3x the algorithmic FLOPs, '
                         'so frac <= 0.333 by construction',
                    by_kind={k: dict(launches=v[0], tflops=(v[1] / (v[2] / 1000.0) / 1e12 if v[2] > 0 else 0.0), ms=v[2]) for k, v in by_kind.items()})

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cres = args.cpu_res or args.res
        v, ph = cpu_reference_iteration(cres, args.cfg, args.cpu_batch)
        cpu = dict(value=v, unit=UNIT, cores=os.cpu_count(), kind='port',
                   sample=f'one amortised iteration (Gmain+Dmain+Greg/4+Dreg/16 incl. Adam) at batch {args.cpu_batch}, {cres}x{cres} '
                          f'{args.cfg} fp32, oracle port of the reference impl=ref path, torch {torch.__version__} CPU, '
                          f'{torch.get_num_threads()} threads',
                   phase_seconds={k: round(x, 3) for k, x in ph.items()})

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_total / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic',
                config=dict(workload=f'StyleGAN2 {args.cfg} (config-f) {args.res}x{args.res} G+D train iteration fp32, batch {args.batch}/GPU '
                                     f'in rounds of {min(args.batch_gpu, args.batch)}', global_batch=args.batch * world,
                            parallelism=f'dp{world}', conv_precision=args.prec,
                            peak_hbm_gb=round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1),
                            l2_policy='inputs and activations (>1 GB per round) exceed the 126 MB L2; no explicit flush',
                            reg_schedule='Greg every 4th, Dreg every 16th iteration, counter reset at the start of the timed region'),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d_bytes), d2h_bytes_per_step=int(d2h_bytes),
                         ms_per_step=ms_e2e / args.steps),
                gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu, clocks=clocks)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
