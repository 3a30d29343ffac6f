#!/usr/bin/env python
"""BASELINE configs[3]: GA StyleSpace-direction population fitness evaluation (64 individuals, shared latent batch of 8,
G.eval / noise_mode='const').  Single process = the per-rank work of `world` ranks is timed as individuals i % world == 0.

    python tools/ga_bench.py [--res 256] [--cfg paper256] [--pop 64] [--batch 8] [--world 1] [--out gpurun_out/ga_bench.txt]
"""
import os, sys, argparse, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gagan_b200  # noqa: E402
_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(_CHECKOUT if os.path.isdir(_CHECKOUT) else None)   # the reference checkout on this build's operators
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--res', type=int, default=256)
    ap.add_argument('--cfg', default='paper256')
    ap.add_argument('--pop', type=int, default=64)
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--world', type=int, default=1)
    ap.add_argument('--out', default='')
    ap.add_argument('--no-graph', action='store_true')
    args = ap.parse_args()
    from torch_utils import custom_ops
    from gagan_b200.training import training_loop, ga_eval
    custom_ops.verbosity = 'none'
    dev = torch.device('cuda:0')
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    G, D = training_loop.build_networks(args.res, args.cfg, device=dev, use_domain_modulation=True, domain_modulation_parametrization='additive')
    pop = ga_eval.init_population(G, args.pop)
    z = torch.randn(args.batch, 512, device=dev)
    lines = []
    for rep in range(3):
        torch.cuda.synchronize()
        l0 = custom_ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        fit = ga_eval.evaluate_population(G, D, pop, z, rank=0, world=args.world, cuda_graph=not args.no_graph)
        e1.record(); torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        n_local = len(ga_eval.shard_indices(args.pop, 0, args.world))
        ms = e0.elapsed_time(e1)
        lines.append(f'{"eager" if args.no_graph else "cuda graph"} rep {rep}: {n_local} individuals x batch {args.batch} @ {args.res}^2 ({args.cfg}): {ms:.1f} ms on the device, {wall * 1e3:.1f} ms wall -> '
                     f'{n_local / ms * 1e3:.1f} individuals/s, {n_local * args.batch / ms * 1e3:.1f} images/s; {custom_ops.launch_count() - l0} library launches; '
                     f'genome {pop.shape[1]}; fitness[0:3] = {[round(float(v), 4) for v in fit[:3]]}')
        print(lines[-1], flush=True)
    if args.out:
        open(args.out, 'a').write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    main()
