#!/usr/bin/env python
"""bench.py -- StyleGAN2 G+D training throughput (images/sec) on B200, BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched per rank by torch.distributed.run)
    python bench.py --impl reference ...                      (the UNMODIFIED reference on its CPU impl='ref' operators)
    python bench.py --workload ada | ga ...                   (BASELINE configs[2] / configs[3]; not the headline)
    python bench.py --global-batch 32 --gpus N                (strong scaling: the reference's `--batch` semantics)

Default workload (config.workload): BASELINE.json configs[1] -- "StyleGAN2 config-f 1024^2 G+D forward-backward fp32,
batch 32 on 1 B200": one step = one training iteration of the upstream loop (Gmain + Dmain every iteration, Greg every
4th, Dreg every 16th, Adam steps, G_ema), batch 32 PER GPU (weak scaling), in `batch_gpu`-sized accumulation rounds exactly
like training_loop.py:495-502.  Synthetic data, random-init weights.  The networks, the loss and (for `ada`) the augment
pipe are the REFERENCE's own classes from the checkout in baseline/_ref (tools/vendor_reference.py), running on this build's
operators through gagan_b200.install() -- the drop-in is what is measured.

One JSON line on stdout (rank 0):
  value      images/sec, inputs resident in HBM when the timed region starts (CUDA events, max over ranks)
  e2e        images/sec through the public API with HOST inputs: every step copies that step's uint8 image batch and
             latents from pinned host memory, and reads the step's loss statistics back (training_stats collector)
  roofline   the dominant kernel family (conv2d fwd/dgrad): algorithmic FLOPs / CUDA-event time of every conv launch inside
             the timed region, against the measured TF32 peak (= 1/2 of the measured bf16 peak)
  cpu_baseline  the reference itself (kind "reference": oracle/_ref through oracle/live_ref.py semantics) timed on this box's
             host cores in a child process (rank 0, N=1)
  fast_mode  the same step with ONE TF32 product per MAC instead of three (not fp32-faithful; reported, never the headline)
"""
import io
import os
import sys
import json
import time
import argparse
import threading
import contextlib
import subprocess

ROOT = os.path.dirname(os.path.abspath(__file__))
CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
ORACLE_REF = os.path.join(ROOT, 'oracle', '_ref', 'DissimilarDomains')
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

METRIC = 'stylegan2_g_d_train_images_per_sec'
UNIT = 'img/s'
AFFINE_PLUS_PARTS = ['synt_affine', 'tRGB_affine', 'synt_weights_offset.b64', 'tRGB_weights_offset.b64']   # DD/README.md:191-196
AFFINE_PLUS_PARAM = 'out_in_additive'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=16)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='train', choices=['train', 'ada', 'ga'],
                    help='train = BASELINE configs[1] (headline); ada = configs[2] (Affine+ few-shot 256^2 with ADA); ga = configs[3]')
    ap.add_argument('--res', type=int, default=None)
    ap.add_argument('--cfg', default=None)
    ap.add_argument('--batch', type=int, default=None, help='images per GPU per iteration (weak scaling)')
    ap.add_argument('--global-batch', type=int, default=None, help='total images per iteration over all GPUs (strong scaling, train.py --batch)')
    ap.add_argument('--batch-gpu', type=int, default=32, help='images per accumulation round (training_loop.py:495-502); one round of 32 fits '
                    "the B200's 180 GB in fp32 and keeps the low-resolution layers' tiles full")
    ap.add_argument('--prec', default='auto', choices=['auto', 'auto_fast', 'simt', 'tf32x1', 'tf32x3'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-fast-mode', action='store_true',
                    help='skip the fast_mode report (4 steps in the tf32x1 mode -- one TF32 product per MAC, NOT fp32-faithful -- timed in a '
                         'CHILD process, so that nothing but the headline workload runs in the process that prints the line)')
    ap.add_argument('--fast-mode', action='store_true', help='(accepted for compatibility: the fast_mode report is on by default)')
    ap.add_argument('--sync-debug', action='store_true', help='debug: synchronise after every library call and name the call that faulted')
    ap.add_argument('--conv-family', type=int, default=1, help='A/B: 0 = the tile kernel serves every tcgen05 convolution (no marching kernel)')
    ap.add_argument('--fused-epilogue', action='store_true', help='A/B: bias / noise / activation in the conv store loop instead of a separate bias_act launch')
    ap.add_argument('--spelled-out-second-order', action='store_true',
                    help='A/B: the round-1 second-order path of the scaled convolution (broadcast multiplies around the unscaled kernels)')
    ap.add_argument('--reference-forwards', action='store_true', help="run the reference's forwards untouched (no fused callers)")
    ap.add_argument('--cpu-res', type=int, default=0, help='resolution of the CPU sample (0 = same as --res)')
    ap.add_argument('--cpu-batch', type=int, default=2)
    ap.add_argument('--population', type=int, default=64)
    ap.add_argument('--latents', type=int, default=8)
    a = ap.parse_args()
    dflt = dict(train=('stylegan2', 1024, 32), ada=('paper256', 256, 8), ga=('paper256', 256, 8))[a.workload]
    a.cfg = a.cfg or dflt[0]
    a.res = a.res or dflt[1]
    world = max(int(os.environ.get('WORLD_SIZE', '1')), 1)
    if a.global_batch is not None:
        assert a.global_batch % world == 0, '--global-batch must be divisible by the number of ranks'
        a.batch = a.global_batch // world
        a.scaling = 'strong'
    else:
        a.batch = a.batch or dflt[2]
        a.scaling = 'weak'
    return a


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


def workload_text(args, world):
    if args.workload == 'train':
        return (f'StyleGAN2 {args.cfg} (config-f) {args.res}x{args.res} G+D train iteration fp32, batch {args.batch}/GPU '
                f'in rounds of {min(args.batch_gpu, args.batch)}')
    if args.workload == 'ada':
        return (f'ADA Affine+ few-shot adaptation {args.res}x{args.res} {args.cfg} fp32 (parts {",".join(AFFINE_PLUS_PARTS)}; {AFFINE_PLUS_PARAM}; '
                f'AugmentPipe bgc, ada_target 0.6; 10 synthetic target images), batch {args.batch}/GPU')
    return (f'GA StyleSpace-direction population fitness eval: {args.population} individuals sharded i % {world}, {args.latents} shared latents, '
            f'{args.res}x{args.res} {args.cfg} G (additive offsets) + D, eval mode, const noise')


# ----------------------------------------------------------------------------------------------------------
# Reference arm / CPU baseline: the UNMODIFIED reference (oracle/_ref, a byte-identical copy of the reference's packages made
# by tools/vendor_reference.py) on CPU tensors, where every op takes its own impl='ref' branch.  Nothing of this build is on
# that path: this function neither imports gagan_b200 nor loads libgagan_b200.so.


REF_CONFIGS = {   # train.py:219-228, for the arm that must not import this build
    'stylegan2': dict(fmaps=1.0, lrate=0.002, gamma=10.0, map=8, mbstd=4), 'paper256': dict(fmaps=0.5, lrate=0.0025, gamma=1.0, map=8, mbstd=8),
    'paper512': dict(fmaps=1.0, lrate=0.0025, gamma=0.5, map=8, mbstd=8), 'paper1024': dict(fmaps=1.0, lrate=0.002, gamma=2.0, map=8, mbstd=4)}
REF_AUGPIPE_BGC = dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1, brightness=1, contrast=1, lumaflip=1, hue=1,
                       saturation=1)   # train.py:365-368


def _load_plain_reference():
    from oracle import live_ref                         # private import + the `img is None` guard applied from outside
    return live_ref.load()


def cpu_reference_iteration(L, args, batch, seed=0):
    """One amortised training iteration of the reference on the host cores; returns (images/sec, seconds per phase)."""
    spec = REF_CONFIGS[args.cfg]
    res = args.cpu_res or args.res
    cb = int(spec['fmaps'] * 32768)
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(seed)
    L.conv2d_gradfix.enabled = True                      # training_loop.py:209-210
    L.grid_sample_gradfix.enabled = True
    extra = {}
    if args.workload == 'ada':
        extra = dict(use_domain_modulation=True, domain_modulation_parametrization=AFFINE_PLUS_PARAM, generator_requires_grad_parts=AFFINE_PLUS_PARTS)
    with contextlib.redirect_stdout(io.StringIO()):
        G = L.networks.Generator(z_dim=512, c_dim=0, w_dim=512, img_resolution=res, img_channels=3, mapping_kwargs=dict(num_layers=spec['map']),
                                 synthesis_kwargs=dict(channel_base=cb, channel_max=512, num_fp16_res=0, conv_clamp=None, **extra)).train()
        D = L.networks.Discriminator(c_dim=0, img_resolution=res, img_channels=3, channel_base=cb, channel_max=512, num_fp16_res=0,
                                     conv_clamp=None, epilogue_kwargs=dict(mbstd_group_size=spec['mbstd'])).train()
    pipe = None
    if args.workload == 'ada':
        pipe = L.augment.AugmentPipe(**REF_AUGPIPE_BGC).train().requires_grad_(False)
        pipe.p.copy_(torch.as_tensor(0.3))
    loss = L.loss.StyleGAN2Loss(device=torch.device('cpu'), G_mapping=G.mapping, G_synthesis=G.synthesis, D=D, augment_pipe=pipe,
                                r1_gamma=spec['gamma'])
    optG = torch.optim.Adam(G.parameters(), lr=spec['lrate'] * 0.8, betas=(0.0, 0.99 ** 0.8), eps=1e-8)
    optD = torch.optim.Adam(D.parameters(), lr=spec['lrate'] * 16 / 17, betas=(0.0, 0.99 ** (16 / 17)), eps=1e-8)
    z = torch.randn(batch, 512)
    c = torch.zeros(batch, 0)
    real = torch.rand(batch, 3, res, res) * 2 - 1
    t = {}
    for name, net, opt, gain in (('Gmain', G, optG, 1), ('Dmain', D, optD, 1), ('Greg', G, optG, 4), ('Dreg', D, optD, 16)):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        G.requires_grad_(False); D.requires_grad_(False)
        if net is G and args.workload == 'ada':
            for n, p in G.named_parameters():           # the Affine+ parts: affine layers + the b64 weight offsets
                p.requires_grad_((('affine' in n) and 'synthesis' in n) or ('weights_offset' in n and 'synthesis.b64' in n))
        else:
            net.requires_grad_(True)
        loss.accumulate_gradients(phase=name, real_img=real, real_c=c, gen_z=z, gen_c=c, sync=True, gain=gain)
        opt.step()
        t[name] = time.perf_counter() - t0
    t_iter = t['Gmain'] + t['Dmain'] + t['Greg'] / 4 + t['Dreg'] / 16
    return batch / t_iter, t


def cpu_reference_ga(L, args, individuals):
    """Fitness of `individuals` GA individuals through the reference G/D on the host cores; returns images/sec."""
    spec = REF_CONFIGS[args.cfg]
    cb = int(spec['fmaps'] * 32768)
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        G = L.networks.Generator(z_dim=512, c_dim=0, w_dim=512, img_resolution=args.res, img_channels=3, mapping_kwargs=dict(num_layers=spec['map']),
                                 synthesis_kwargs=dict(channel_base=cb, channel_max=512, num_fp16_res=0, conv_clamp=None, use_domain_modulation=True,
                                                       domain_modulation_parametrization='additive')).eval()
        D = L.networks.Discriminator(c_dim=0, img_resolution=args.res, img_channels=3, channel_base=cb, channel_max=512, num_fp16_res=0,
                                     conv_clamp=None, epilogue_kwargs=dict(mbstd_group_size=spec['mbstd'])).eval()
    z = torch.randn(args.latents, 512); c = torch.zeros(args.latents, 0)
    layers = [m for m in G.synthesis.modules() if isinstance(getattr(m, 'offset', None), torch.nn.Parameter)]
    t0 = time.perf_counter()
    with torch.no_grad():
        ws = G.mapping(z, c)
        for _ in range(individuals):
            for m in layers:
                m.offset.copy_(0.1 * torch.randn(m.offset.shape))
            D(G.synthesis(ws, noise_mode='const'), c).mean().item()
    return individuals * args.latents / (time.perf_counter() - t0)


def run_reference_arm(args, rank):
    """`--impl reference`: the reference itself on all host threads, rank 0 only."""
    if rank != 0:
        return
    if not os.path.isdir(os.path.join(ORACLE_REF, 'training')):
        print(json.dumps(dict(impl='reference', unavailable='oracle/_ref is missing (tools/vendor_reference.py needs /root/reference)')), flush=True)
        return
    L = _load_plain_reference()
    res = args.cpu_res or args.res
    budget_s = 240.0
    t_start = time.perf_counter()
    vals, phases, done_warm, done = [], None, 0, 0
    if args.workload == 'ga':
        n_ind = 2
        vals = [cpu_reference_ga(L, args, n_ind)]
        done, phases = 1, {}
        sample = f'{n_ind} individuals x {args.latents} latents at {args.res}x{args.res} {args.cfg}, reference G/D eval on the host cores'
        batch_for_ms = n_ind * args.latents
    else:
        for i in range(args.warmup + args.steps):
            v, ph = cpu_reference_iteration(L, args, args.cpu_batch, seed=i)
            if i >= args.warmup or (time.perf_counter() - t_start) > budget_s * 0.5:
                vals.append(v); phases = ph; done += 1
            else:
                done_warm += 1
            if (time.perf_counter() - t_start) > budget_s and vals:
                break
        sample = (f'{done} timed + {done_warm} warm-up amortised iterations (Gmain+Dmain+Greg/4+Dreg/16 incl. Adam) at batch '
                  f'{args.cpu_batch}, {res}x{res} {args.cfg} fp32, the unmodified reference (training/networks.py, training/loss.py on its '
                  f'impl=ref operators), torch {torch.__version__} CPU, {torch.get_num_threads()} threads; wall budget {budget_s:.0f}s')
        batch_for_ms = args.cpu_batch
    value = float(np.mean(vals))
    world = max(args.gpus, 1)
    line = dict(metric=METRIC if args.workload != 'ga' else 'ga_population_fitness_images_per_sec', value=value, unit=UNIT, impl='reference',
                n_gpus=args.gpus, steps=done, warmup=done_warm, ms_per_step=1000.0 * batch_for_ms / value, higher_is_better=True,
                scaling=args.scaling, vs_baseline=None, dtype='f32', data='synthetic',
                # the same workload as the GPU arm (its `config.workload`, `global_batch`); what was actually timed is `cpu_baseline.sample`
                config=dict(workload=workload_text(args, world), global_batch=args.batch * world, parallelism=f'dp{world}',
                            reference_sample=f'batch {args.cpu_batch} at {res}x{res} on the host cores, scaled per image'),
                cpu_baseline=dict(value=value, unit=UNIT, cores=os.cpu_count(), kind='reference', sample=sample,
                                  phase_seconds={k: round(v, 3) for k, v in (phases or {}).items()}),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def cpu_baseline_child(args):
    """The cpu_baseline leg of the GPU arm: the reference arm as a child process (the parent has this build's operators installed
    under the reference's module names, the child is a clean interpreter that never loads them)."""
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '1', '--warmup', '0', '--workload', args.workload,
           '--cfg', args.cfg, '--res', str(args.res), '--cpu-res', str(args.cpu_res), '--cpu-batch', str(args.cpu_batch),
           '--batch', str(args.batch), '--population', str(args.population), '--latents', str(args.latents)]
    env = dict(os.environ, WORLD_SIZE='1', RANK='0', LOCAL_RANK='0', CUDA_VISIBLE_DEVICES='')
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900, env=env)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith('{'):
                d = json.loads(ln)
                return d.get('cpu_baseline', d)
        return dict(unavailable=(out.stderr or 'no output')[-300:])
    except Exception as e:       # a failed baseline must not lose the GPU measurement
        return dict(unavailable=str(e)[-300:])


def fast_mode_child(args):
    """The fast_mode report: the same step with ONE TF32 product per MAC (conv_precision = PREC_AUTO_FAST), 4 timed steps in a child
    process.  Not fp32-faithful and never the headline; a failure here cannot touch the measurement of the parent."""
    cmd = [sys.executable, os.path.abspath(__file__), '--prec', 'auto_fast', '--steps', '4', '--warmup', '2', '--no-cpu-baseline', '--no-fast-mode',
           '--workload', 'train', '--cfg', args.cfg, '--res', str(args.res), '--batch', str(args.batch), '--conv-family', str(args.conv_family)]
    env = dict(os.environ, WORLD_SIZE='1', RANK='0', LOCAL_RANK='0')
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith('{'):
                d = json.loads(ln)
                return dict(mode='tf32x1 (one TF32 product per MAC instead of hi*hi + hi*lo + lo*hi), timed in a child process', value=d['value'],
                            unit=d['unit'], steps=d['steps'], ms_per_step=d['ms_per_step'],
                            parity='NOT fp32-faithful: network-level max-rel-err of the four loss-phase goldens and the config-size layers is '
                                   'recorded by tests/test_gpu_config_size.py::test_fast_mode_report -> profiles/r2_fast_mode_parity.txt; '
                                   'the headline stays the 3xTF32 mode')
        return dict(unavailable=(out.stderr or 'no output')[-300:])
    except Exception as e:
        return dict(unavailable=str(e)[-300:])


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
    import gagan_b200
    if not os.path.isdir(os.path.join(CHECKOUT, 'training')):
        raise SystemExit('bench.py: baseline/_ref/DissimilarDomains is missing -- the bench drives the reference checkout on this build; '
                         'run `python tools/vendor_reference.py` where /root/reference exists')
    gagan_b200.install(CHECKOUT, fused_callers=not args.reference_forwards)
    from torch_utils import custom_ops
    from gagan_b200.training import training_loop, ga_eval

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
    PREC = dict(auto=custom_ops.PREC_AUTO, auto_fast=custom_ops.PREC_AUTO_FAST, simt=custom_ops.PREC_FP32_SIMT, tf32x1=custom_ops.PREC_TF32X1, tf32x3=custom_ops.PREC_TF32X3)
    custom_ops.conv_precision = PREC[args.prec]
    custom_ops.set_conv_kernel_family(args.conv_family)
    if args.sync_debug:
        plug = custom_ops.get_plugin('conv2d_plugin')
        for name in ('conv2d', 'conv2d_wgrad', 'bias_act', 'bias_act_noise', 'upfirdn2d', 'fir4_pm', 'chan_dot'):
            def wrap(orig, name=name):
                def fn(*a, **kw):
                    out = orig(*a, **kw)
                    try:
                        torch.cuda.current_stream().synchronize()
                    except Exception as e:
                        shapes = [tuple(t.shape) for t in a if isinstance(t, torch.Tensor)]
                        print(f'[rank {rank}] {custom_ops.watchdog_report()}', file=sys.stderr, flush=True)
                        print(f'[rank {rank}] FAULT after {name} {shapes} { {k: (tuple(v.shape) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()} }: {e}', file=sys.stderr, flush=True)
                        raise
                    return out
                return fn
            setattr(plug, name, wrap(getattr(plug, name)))
    if args.fused_epilogue:
        from torch_utils.ops import conv2d_gradfix
        conv2d_gradfix.fuse_epilogue = True
    if args.spelled_out_second_order:
        from torch_utils.ops import conv2d_gradfix
        conv2d_gradfix.closed_scaled_backward = False

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    spec = training_loop.CONFIGS[args.cfg]
    torch.manual_seed(0 * world + rank)                       # training_loop.py:204-205

    if args.workload == 'ga':
        G, D = training_loop.build_networks(args.res, args.cfg, device=dev, use_domain_modulation=True, domain_modulation_parametrization='additive')
        pop = ga_eval.init_population(G, args.population, seed=0)
        z_host = torch.randn(args.latents, 512, generator=torch.Generator().manual_seed(1)).pin_memory()
        z_dev = z_host.to(dev)
        units = args.population * args.latents               # images through G and D per step, whole job

        def one_step(host_inputs):
            z = z_host.to(dev, non_blocking=True) if host_inputs else z_dev
            fit = ga_eval.evaluate_population(G, D, pop, z, rank=rank, world=world)
            return fit.cpu() if host_inputs else fit
        h2d_bytes, metric = z_host.numel() * 4 + pop.numel() * 4, 'ga_population_fitness_images_per_sec'
    else:
        extra, step_kw = {}, {}
        if args.workload == 'ada':
            extra = dict(use_domain_modulation=True, domain_modulation_parametrization=AFFINE_PLUS_PARAM,
                         generator_requires_grad_parts=AFFINE_PLUS_PARTS)
            step_kw = dict(g_parts=AFFINE_PLUS_PARTS, glrate=0.02, augment_kwargs=training_loop.AUGPIPE_BGC, ada_target=0.6, augment_p=0.0)
        G, D = training_loop.build_networks(args.res, args.cfg, device=dev, **extra)
        step = training_loop.TrainingStep(G, D, batch_size=args.batch * world, batch_gpu=min(args.batch_gpu, args.batch), device=dev,
                                          lrate=spec['lrate'], r1_gamma=spec['gamma'], ema_kimg=spec['ema'], rank=rank, num_gpus=world, **step_kw)
        n_phases = len(step.phases)
        # synthetic inputs: device-resident for `value`, pinned host uint8 + host latents for `e2e`
        n_real = args.batch if args.workload == 'train' else 10       # ada: a 10-shot target set cycled by the sampler
        real_host_set = torch.randint(0, 256, (n_real, 3, args.res, args.res), dtype=torch.uint8)
        real_host = real_host_set[torch.arange(args.batch) % n_real].contiguous().pin_memory()
        real_dev = real_host.to(dev).to(torch.float32) / 127.5 - 1    # training_loop.py:441
        z_host = torch.randn(n_phases, args.batch, 512).pin_memory()
        units = args.batch * world
        h2d_bytes, metric = real_host.numel() + z_host.numel() * 4, METRIC

        def one_step(host_inputs):
            if host_inputs:
                real = real_host.to(dev, non_blocking=True).to(torch.float32) / 127.5 - 1
                step.run(real, z_host.to(dev, non_blocking=True))
                return step.read_stats()                  # the step's loss statistics -> host (syncs; all-reduce across ranks)
            step.run(real_dev)
            return None

    def timed(n_steps, host_inputs):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.workload != 'ga':
            step.cur_it = 0                       # Greg fires ceil(K/4) times, Dreg ceil(K/16) times: never under-counted
        out = None
        e0.record()
        for _ in range(n_steps):
            out = one_step(host_inputs)
        e1.record()
        barrier()
        d2h = 0
        if host_inputs and out is not None:
            d2h = (out.numel() * 4) if isinstance(out, torch.Tensor) else 3 * 8 * len(out)     # training_stats moves 3 fp64 moments per name
        return max_over_ranks(e0.elapsed_time(e1)), d2h

    for _ in range(max(args.warmup, 0)):                      # warm-up covers all four phases (cur_it = 0 fires both regs)
        if args.workload != 'ga':
            step.cur_it = 0
        one_step(False)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    custom_ops.conv_profile = [] if args.workload != 'ga' else None     # (the GA evaluation replays CUDA graphs: no events inside a capture)
    launches0 = custom_ops.launch_count()
    ms_total, _ = timed(args.steps, host_inputs=False)
    launches = custom_ops.launch_count() - launches0
    prof, custom_ops.conv_profile = (custom_ops.conv_profile or []), None
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, d2h_bytes = timed(args.steps, host_inputs=True)

    fast = None        # filled in by rank 0 after the timed regions (fast_mode_child)

    value = units * args.steps / (ms_total / 1000.0)
    e2e_value = units * args.steps / (ms_e2e / 1000.0)

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
    # DRAM bytes per launch of the same kernel family from the committed ncu capture of this command (profiles/): bench.py cannot run
    # a profiler itself, so the figure is read from the summary that tools/ncu_traffic.py wrote; null if there is none
    traffic, traffic_note = None, 'no ncu capture committed'
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.isfile(tpath):
        try:
            tj = json.load(open(tpath))
            traffic, traffic_note = tj['conv_tc_kernel']['dram_bytes_per_launch'], tj['conv_tc_kernel']['note']
        except Exception as e:      # a malformed summary must not break the bench line
            traffic_note = f'unreadable {tpath}: {e}'
    roofline = dict(bound='tensor', achieved=(achieved if prof else None), peak=tf32_peak, unit='TFLOP/s', frac=(achieved / tf32_peak if prof else None), traffic=traffic,
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
        cpu = cpu_baseline_child(args)
    if world == 1 and not args.no_fast_mode and args.prec == 'auto' and args.workload == 'train':
        torch.cuda.empty_cache()                  # the child needs the same ~64 GB next to this process
        fast = fast_mode_child(args)

    cfg = dict(workload=workload_text(args, world), global_batch=args.batch * world, parallelism=f'dp{world}', conv_precision=args.prec,
               callers=f'reference checkout (training/networks.py, training/loss.py) + gagan_b200.install(fused_callers={not args.reference_forwards})',
               peak_hbm_gb=round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1),
               l2_policy='inputs and activations (>1 GB per round) exceed the 126 MB L2; no explicit flush')
    if args.workload != 'ga':
        cfg['reg_schedule'] = 'Greg every 4th, Dreg every 16th iteration, counter reset at the start of the timed region'
    else:
        cfg['individuals_per_sec'] = args.population * args.steps / (ms_total / 1000.0)
    line = dict(metric=metric, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_total / args.steps, higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype='f32',
                data='synthetic', config=cfg,
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d_bytes), d2h_bytes_per_step=int(d2h_bytes),
                         ms_per_step=ms_e2e / args.steps),
                gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu, fast_mode=fast, clocks=clocks)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    try:
        main()
    except BaseException as e:
        if not isinstance(e, SystemExit):
            try:                                   # readable even after the CUDA context faulted (host-mapped record)
                import gagan_b200  # noqa: F401
                from gagan_b200.torch_utils import custom_ops as _co
                rep = _co.watchdog_report()
                if rep:
                    print(f'[rank {os.environ.get("RANK", "0")}] {rep}', file=sys.stderr, flush=True)
            except Exception:
                pass
        raise
