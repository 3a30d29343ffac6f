"""GPU, 2 devices, NCCL: the minibatch sharding of the training step on the REAL networks.

Two ranks, each with half of the batch and the modules wrapped in DistributedDataParallel where the reference wraps them
(training_loop.py:270-285), must end a loss phase with the SAME gradients as one process that saw the whole batch -- the
all-reduce (mean) of NCCL over NVLink is the only collective on this path.  Also the GA population evaluation: individuals
i % world, one all_gather, identical fitness vectors on both ranks and equal to the single-process result.

Skipped on a box with fewer than two GPUs (the driver's default `-m gpu` run uses one); run with `gpurun --gpus 2`.
"""
import os
import socket
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import HAVE_CHECKOUT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(networks, quiet, dev, ga=False):
    extra = dict(use_domain_modulation=True, domain_modulation_parametrization='additive') if ga else {}
    torch.manual_seed(0)
    G = quiet(networks.Generator, z_dim=64, c_dim=0, w_dim=64, img_resolution=64, img_channels=3, mapping_kwargs=dict(num_layers=2),
              synthesis_kwargs=dict(channel_base=2048, channel_max=64, **extra))
    D = quiet(networks.Discriminator, c_dim=0, img_resolution=64, img_channels=3, channel_base=2048, channel_max=64,
              epilogue_kwargs=dict(mbstd_num_channels=0))     # minibatch-std statistics are per-GPU by design (networks.py:1284-1301)
    return G.to(dev), D.to(dev)


def _worker(rank, world, port, out_dir):
    import tests.util as U                                   # installs the drop-in in this process
    from training import networks, loss as loss_mod
    from gagan_b200.training import ga_eval
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    G, D = _build(networks, U.quiet, dev)
    G.mapping.w_avg_beta = None
    gen = torch.Generator().manual_seed(7)
    batch = 8
    real = (torch.rand(batch, 3, 64, 64, generator=gen) * 2 - 1).to(dev)
    z = torch.randn(batch, 64, generator=gen).to(dev)
    c = torch.zeros(batch, 0, device=dev)

    def phase_grads(G_map, G_syn, D_, sl, phase, net):
        for m in (G, D):
            m.requires_grad_(False)
            for p in m.parameters():
                p.grad = None
        net.requires_grad_(True)
        L = loss_mod.StyleGAN2Loss(device=dev, G_mapping=G_map, G_synthesis=G_syn, D=D_, style_mixing_prob=0, r1_gamma=1.0, pl_weight=2.0)
        L.accumulate_gradients(phase=phase, real_img=real[sl], real_c=c[sl], gen_z=z[sl], gen_c=c[sl], sync=True, gain=1)
        return {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None and 'noise_strength' not in k}

    full = {}
    for phase, net in (('Dmain', D), ('Dreg', D), ('Gmain', G)):
        full[phase] = phase_grads(G.mapping, G.synthesis, D, slice(0, batch), phase, net)      # one process, whole batch
    ddp = {}
    for name, module in (('G_mapping', G.mapping), ('G_synthesis', G.synthesis), ('D', D)):
        module.requires_grad_(True)
        ddp[name] = torch.nn.parallel.DistributedDataParallel(module, device_ids=[dev], broadcast_buffers=False)
        module.requires_grad_(False)
    per = batch // world
    sl = slice(rank * per, (rank + 1) * per)
    worst = {}
    for phase, net in (('Dmain', D), ('Dreg', D), ('Gmain', G)):
        got = phase_grads(ddp['G_mapping'], ddp['G_synthesis'], ddp['D'], sl, phase, net)
        w = 0.0
        for k, g in got.items():
            want = full[phase][k]
            denom = float(want.abs().max())
            if denom > 0:
                w = max(w, float((g - want).abs().max()) / denom)
        worst[phase] = w

    # GA population evaluation sharded by individual
    Gg, Dg = _build(networks, U.quiet, dev, ga=True)
    pop = ga_eval.init_population(Gg, size=7, scale=0.1, seed=5)
    zz = torch.randn(4, 64, generator=torch.Generator().manual_seed(9)).to(dev)
    fit = ga_eval.evaluate_population(Gg, Dg, pop, zz, rank=rank, world=world)
    fit1 = ga_eval.evaluate_population(Gg, Dg, pop, zz, rank=0, world=1)
    # two GA generations: the new population is drawn on rank 0 and broadcast over NCCL (ga_eval.next_generation)
    pop2, hist = ga_eval.evolve(Gg, Dg, pop, zz, generations=2, rank=rank, world=world, seed=3)
    torch.save(dict(worst=worst, fit=fit.cpu(), fit1=fit1.cpu(), pop2=pop2.cpu(), hist=[h.cpu() for h in hist]), os.path.join(out_dir, f'r{rank}.pt'))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_gradients_and_ga_sharding_over_nccl():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
    if not HAVE_CHECKOUT:
        pytest.skip('baseline/_ref/DissimilarDomains is absent')
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, _free_port(), d), nprocs=2, join=True)
        r0 = torch.load(os.path.join(d, 'r0.pt')); r1 = torch.load(os.path.join(d, 'r1.pt'))
    print('DDP(2 x 4 images) vs one process (8 images), worst max-rel-err per phase:', r0['worst'], r1['worst'])
    for r in (r0, r1):
        for phase, w in r['worst'].items():
            assert w <= 2e-4, (phase, w)
    assert torch.equal(r0['fit'], r1['fit'])
    assert torch.allclose(r0['fit'], r0['fit1'], rtol=1e-5, atol=1e-6)
    assert torch.equal(r0['pop2'], r1['pop2']) and all(torch.equal(a, b) for a, b in zip(r0['hist'], r1['hist']))
    assert float(r0['hist'][1].max()) >= float(r0['hist'][0].max()) - 1e-6            # elitism
