"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic:

  * the G+D training step shards by minibatch: `TrainingStep` wraps the modules in DistributedDataParallel exactly where
    the reference does (training_loop.py:270-285) and all-reduces the phase's gradients once, on its last accumulation
    round -- two ranks with half the batch each must end with the SAME parameters as one process with the whole batch;
  * GA population evaluation shards by individual (i % world) with ONE all_gather of the fitness slices.

The CUDA ops have no CPU path, so the first two tests run stand-in networks with the same module interface (mapping /
synthesis / D(img, c)): under test are the partitioning, the sync flags and the collectives.  The third runs the REFERENCE's
networks and loss on this build's host layer with the kernels replaced by the torch stand-in (tests/fake_plugin.py): the CPU
twin of tests/test_gpu_multigpu.py.
"""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import tests.util  # noqa: F401  (sys.path)


class _Mapping(torch.nn.Module):
    def __init__(self, z_dim, w_dim, num_ws):
        super().__init__()
        self.fc = torch.nn.Linear(z_dim, w_dim)
        self.num_ws = num_ws

    def forward(self, z, c, skip_w_avg_update=False, **_):
        return torch.tanh(self.fc(z)).unsqueeze(1).repeat(1, self.num_ws, 1)


class _Layer(torch.nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.offset = torch.nn.Parameter(torch.zeros([1, ch]))


class _Synthesis(torch.nn.Module):
    def __init__(self, w_dim, res):
        super().__init__()
        self.res = res
        self.l0 = _Layer(w_dim)
        self.l1 = _Layer(w_dim)
        self.fc = torch.nn.Linear(w_dim, 3 * res * res)

    def forward(self, ws, noise_mode='random', **_):
        s = ws[:, 0] + self.l0.offset + 2.0 * self.l1.offset
        return torch.tanh(self.fc(s)).reshape(-1, 3, self.res, self.res)


class _G(torch.nn.Module):
    def __init__(self, z_dim=8, w_dim=8, res=4):
        super().__init__()
        self.z_dim, self.w_dim = z_dim, w_dim
        self.mapping = _Mapping(z_dim, w_dim, 2)
        self.synthesis = _Synthesis(w_dim, res)

    def forward(self, z, c, **kw):
        return self.synthesis(self.mapping(z, c), **kw)


class _D(torch.nn.Module):
    def __init__(self, res=4):
        super().__init__()
        self.fc0 = torch.nn.Linear(3 * res * res, 16)
        self.fc1 = torch.nn.Linear(16, 1)

    def forward(self, img, c, **_):
        return self.fc1(torch.nn.functional.softplus(self.fc0(img.flatten(1))))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run_step(rank, world, port, out_dir):
    from gagan_b200.training.training_loop import TrainingStep
    if world > 1:
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)                                   # identical initial weights on every rank
    G, D = _G(), _D()
    gen = torch.Generator().manual_seed(123)
    batch = 8
    real = torch.rand(batch, 3, 4, 4, generator=gen) * 2 - 1
    step = TrainingStep(G, D, batch_size=batch, batch_gpu=2, device='cpu', lrate=0.01, r1_gamma=0.0, pl_weight=0.0,
                        style_mixing_prob=0.0, rank=rank, num_gpus=world)
    zs = torch.randn(len(step.phases), batch, 8, generator=gen)
    per = batch // world
    sl = slice(rank * per, (rank + 1) * per)
    # count gradient all-reduces: DDP must sync once per phase with a backward pass (last accumulation round only)
    syncs = []
    orig = step.loss.accumulate_gradients

    def counted(phase, real_img, real_c, gen_z, gen_c, sync, gain):
        syncs.append((phase, bool(sync)))
        return orig(phase=phase, real_img=real_img, real_c=real_c, gen_z=gen_z, gen_c=gen_c, sync=sync, gain=gain)
    step.loss.accumulate_gradients = counted
    for _ in range(2):
        step.run(real[sl], zs[:, sl])
    params = {('G.' + k): v.detach().clone() for k, v in G.state_dict().items()}
    params.update({('D.' + k): v.detach().clone() for k, v in D.state_dict().items()})
    params.update({('Gema.' + k): v.detach().clone() for k, v in step.G_ema.state_dict().items()})
    torch.save(dict(params=params, syncs=syncs), os.path.join(out_dir, f'step_w{world}_r{rank}.pt'))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_training_step_shards_by_minibatch_and_allreduces_once_per_phase():
    with tempfile.TemporaryDirectory() as d:
        _run_step(0, 1, 0, d)
        mp.spawn(_run_step, args=(2, _free_port(), d), nprocs=2, join=True)
        one = torch.load(os.path.join(d, 'step_w1_r0.pt'))
        r0 = torch.load(os.path.join(d, 'step_w2_r0.pt'))
        r1 = torch.load(os.path.join(d, 'step_w2_r1.pt'))
    for k, v in one['params'].items():
        assert torch.allclose(r0['params'][k], r1['params'][k], rtol=0, atol=0), f'ranks diverged on {k}'
        # DDP averages over ranks what one process sums over twice as many rounds (factor 2 on the gradients, exactly as in
        # the reference loop); Adam is scale-invariant up to its eps, hence the small absolute tolerance
        assert torch.allclose(r0['params'][k], v, rtol=1e-4, atol=2e-5), f'2-rank result differs from 1 process on {k}'
    # 8 images / (2 per round * 2 ranks) = 2 rounds per phase: sync only on the last one (training_loop.py:498)
    by_phase = {}
    for phase, sync in r0['syncs']:
        by_phase.setdefault(phase, []).append(sync)
    for phase, flags in by_phase.items():
        assert flags == [False, True] * (len(flags) // 2), (phase, flags)
    # the single process runs 4 rounds per phase
    assert len(one['syncs']) == 2 * len(r0['syncs'])


def _run_ga(rank, world, port, out_dir):
    from gagan_b200.training import ga_eval
    if world > 1:
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    G, D = _G(), _D()
    pop = ga_eval.init_population(G, size=7, scale=0.1, seed=5)         # 7 individuals: ragged over 2 ranks (4 + 3)
    assert pop.shape == (7, ga_eval.genome_size(G)) and ga_eval.genome_size(G) == 16
    z = torch.randn(4, 8, generator=torch.Generator().manual_seed(9))
    seen = []

    def fitness(G_, D_, ws, c):
        seen.append(float(G_.synthesis.l0.offset.sum() + G_.synthesis.l1.offset.sum()))
        return ga_eval.default_fitness(G_, D_, ws, c)
    fit = ga_eval.evaluate_population(G, D, pop, z, rank=rank, world=world, fitness_fn=fitness)
    torch.save(dict(fit=fit, seen=seen, idx=ga_eval.shard_indices(7, rank, world)), os.path.join(out_dir, f'ga_w{world}_r{rank}.pt'))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_ga_population_eval_shards_by_individual_with_one_all_gather():
    with tempfile.TemporaryDirectory() as d:
        _run_ga(0, 1, 0, d)
        mp.spawn(_run_ga, args=(2, _free_port(), d), nprocs=2, join=True)
        one = torch.load(os.path.join(d, 'ga_w1_r0.pt'))
        r0 = torch.load(os.path.join(d, 'ga_w2_r0.pt'))
        r1 = torch.load(os.path.join(d, 'ga_w2_r1.pt'))
    assert r0['idx'] == [0, 2, 4, 6] and r1['idx'] == [1, 3, 5]
    assert len(r0['seen']) == 4 and len(r1['seen']) == 3                  # each rank evaluated only its own individuals
    assert torch.equal(r0['fit'], r1['fit'])                              # every rank holds the full fitness vector
    assert torch.allclose(r0['fit'], one['fit'], rtol=1e-6, atol=1e-7)
    assert not torch.isnan(r0['fit']).any()
    assert len(set(np.round(one['fit'].numpy(), 6))) > 1                  # the offsets do change the fitness


# ----------------------------------------------------------------------------
# The same two workloads on the REFERENCE's networks (the installed checkout) over this build's host layer, kernels replaced by the
# torch stand-in: the CPU twin of tests/test_gpu_multigpu.py (2 NCCL ranks on 2 GPUs).

def _host_layer_on_cpu():
    from tests.fake_plugin import FakePlugin
    from torch_utils.ops import conv2d_gradfix as cg, bias_act as BA, upfirdn2d as U
    from torch_utils import custom_ops
    fp = FakePlugin()
    cg._plugin = fp; BA._plugin = fp; U._plugin = fp
    for name in ('bias_act_plugin', 'upfirdn2d_plugin', 'conv2d_plugin'):
        custom_ops._cached_plugins[name] = fp
    for mod in (cg, U, BA):
        mod._check_input = lambda t: None


def _run_real_networks(rank, world, port, out_dir):
    import tests.util as U_                                   # installs the drop-in in this process
    _host_layer_on_cpu()
    from training import networks, loss as loss_mod
    from gagan_b200.training import ga_eval
    torch.set_num_threads(2)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    dev = torch.device('cpu')

    def build(ga=False):
        extra = dict(use_domain_modulation=True, domain_modulation_parametrization='additive') if ga else {}
        torch.manual_seed(0)
        G = U_.quiet(networks.Generator, z_dim=16, c_dim=0, w_dim=16, img_resolution=16, img_channels=3, mapping_kwargs=dict(num_layers=2),
                     synthesis_kwargs=dict(channel_base=256, channel_max=16, **extra))
        D = U_.quiet(networks.Discriminator, c_dim=0, img_resolution=16, img_channels=3, channel_base=256, channel_max=16,
                     epilogue_kwargs=dict(mbstd_num_channels=0))     # minibatch-std statistics are per-rank by design (networks.py:1284-1301)
        return G, D
    G, D = build()
    G.mapping.w_avg_beta = None
    gen = torch.Generator().manual_seed(7)
    batch = 8
    real = torch.rand(batch, 3, 16, 16, generator=gen) * 2 - 1
    z = torch.randn(batch, 16, generator=gen)
    c = torch.zeros(batch, 0)

    def phase_grads(G_map, G_syn, D_, sl, phase, net):
        for m in (G, D):
            m.requires_grad_(False)
            for p in m.parameters():
                p.grad = None
        net.requires_grad_(True)
        L = loss_mod.StyleGAN2Loss(device=dev, G_mapping=G_map, G_synthesis=G_syn, D=D_, style_mixing_prob=0, r1_gamma=1.0, pl_weight=2.0)
        L.accumulate_gradients(phase=phase, real_img=real[sl], real_c=c[sl], gen_z=z[sl], gen_c=c[sl], sync=True, gain=1)
        return {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None and 'noise_strength' not in k}

    phases = (('Dmain', D), ('Dreg', D), ('Gmain', G))
    full = {phase: phase_grads(G.mapping, G.synthesis, D, slice(0, batch), phase, net) for phase, net in phases}
    ddp = {}
    for name, module in (('G_mapping', G.mapping), ('G_synthesis', G.synthesis), ('D', D)):
        module.requires_grad_(True)
        ddp[name] = torch.nn.parallel.DistributedDataParallel(module, broadcast_buffers=False)      # training_loop.py:270-285
        module.requires_grad_(False)
    per = batch // world
    sl = slice(rank * per, (rank + 1) * per)
    worst = {}
    for phase, net in phases:
        got = phase_grads(ddp['G_mapping'], ddp['G_synthesis'], ddp['D'], sl, phase, net)
        assert set(got) == set(full[phase])
        w = 0.0
        for k, g in got.items():
            want = full[phase][k]
            denom = float(want.abs().max())
            if denom > 0:
                w = max(w, float((g - want).abs().max()) / denom)
        worst[phase] = w

    Gg, Dg = build(ga=True)
    pop = ga_eval.init_population(Gg, size=7, scale=0.1, seed=5)
    zz = torch.randn(4, 16, generator=torch.Generator().manual_seed(9))
    fit = ga_eval.evaluate_population(Gg, Dg, pop, zz, rank=rank, world=world)
    fit1 = ga_eval.evaluate_population(Gg, Dg, pop, zz, rank=0, world=1)
    torch.save(dict(worst=worst, fit=fit, fit1=fit1), os.path.join(out_dir, f'real_r{rank}.pt'))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_gradients_and_ga_sharding_on_the_reference_networks_over_gloo():
    """Two gloo ranks with half of the batch each, the reference's mapping / synthesis / discriminator wrapped in
    DistributedDataParallel where the reference wraps them, end a loss phase (Dmain, Dreg with its double backward, Gmain) with the
    gradients of one process that saw the whole batch; the GA population evaluation gives every rank the single-process fitness
    vector.  Host layer of this build, stand-in kernels."""
    if not tests.util.HAVE_CHECKOUT:
        pytest.skip('baseline/_ref/DissimilarDomains is absent')
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_run_real_networks, args=(2, _free_port(), d), nprocs=2, join=True)
        r0 = torch.load(os.path.join(d, 'real_r0.pt')); r1 = torch.load(os.path.join(d, 'real_r1.pt'))
    for r in (r0, r1):
        for phase, w in r['worst'].items():
            assert w <= 2e-4, (phase, w)
    assert torch.equal(r0['fit'], r1['fit'])
    assert torch.allclose(r0['fit'], r0['fit1'], rtol=1e-5, atol=1e-6)


# ----------------------------------------------------------------------------
# One GA generation: selection on the gathered fitness, the reference's operators, one broadcast of the new population.

def test_ga_operators_match_the_reference_on_shared_draws():
    """gaussian_crossover / dynamic_mutation of ga_eval against GA/crossover_mutation.py:4-7,16-19 with the random draws routed
    through one seeded generator on both sides; next_generation keeps the elite, breeds from the fitter half and ranks NaN last."""
    from oracle import live_ref
    from gagan_b200.training import ga_eval
    ref = live_ref.load_ga_operators()
    if ref is None:
        pytest.skip('oracle/_ref/GA is absent')
    g = torch.Generator().manual_seed(1)
    p1, p2 = torch.randn(5, 33, generator=g), torch.randn(5, 33, generator=g)
    with tests.util.patched_randn(4):
        want_c = ref.gaussian_crossover(p1, p2)
        want_m = ref.dynamic_mutation(want_c, mutation_rate=0.07)
    with tests.util.patched_randn(4):
        got_c = ga_eval.gaussian_crossover(p1, p2)
        got_m = ga_eval.dynamic_mutation(got_c, mutation_rate=0.07)
    assert torch.equal(got_c, want_c) and torch.equal(got_m, want_m)

    pop = torch.randn(8, 33, generator=g)
    fit = torch.tensor([0.3, float('nan'), 2.0, -1.0, 0.9, 0.1, 1.5, -0.2])
    new = ga_eval.next_generation(pop, fit, elite=2, mutation_rate=0.0, seed=3)
    assert torch.equal(new[0], pop[2]) and torch.equal(new[1], pop[6])                 # the two fittest survive unchanged, in order
    assert torch.equal(new, ga_eval.next_generation(pop, fit, elite=2, mutation_rate=0.0, seed=3))        # reproducible
    assert not torch.equal(new[2:], ga_eval.next_generation(pop, fit, elite=2, mutation_rate=0.0, seed=4)[2:])
    # the children are bred from the fitter half {2, 6, 4, 0} only: changing the genomes of the weaker half (and of the NaN
    # individual) changes nothing
    other = pop.clone()
    other[[1, 3, 5, 7]] = torch.randn(4, 33, generator=g)
    assert torch.equal(ga_eval.next_generation(other, fit, elite=2, mutation_rate=0.0, seed=3), new)
    assert torch.isfinite(new).all()
    # mutation adds mutation_rate * N(0,1) on top of the same children
    mut = ga_eval.next_generation(pop, fit, elite=2, mutation_rate=0.05, seed=3)
    assert torch.equal(mut[:2], new[:2]) and 0.02 < float((mut[2:] - new[2:]).std()) < 0.1


def _run_generation(rank, world, port, out_dir):
    from gagan_b200.training import ga_eval
    if world > 1:
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                                          # the ranks' own RNG streams differ on purpose
    G, D = _G(), _D()
    G.load_state_dict(torch.load(os.path.join(out_dir, 'G0.pt'))); D.load_state_dict(torch.load(os.path.join(out_dir, 'D0.pt')))
    pop = ga_eval.init_population(G, size=7, scale=0.1, seed=5)
    z = torch.randn(4, 8, generator=torch.Generator().manual_seed(9))
    final, history = ga_eval.evolve(G, D, pop, z, generations=3, rank=rank, world=world, seed=11, elite=2)
    torch.save(dict(final=final, history=history), os.path.join(out_dir, f'gen_w{world}_r{rank}.pt'))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_ga_generations_broadcast_one_population_to_every_rank():
    """Three generations (evaluate -> select -> crossover -> mutate -> broadcast) on two gloo ranks whose own RNG streams differ: both
    ranks hold the same population after every generation, equal to the single-process run (the draws happen on rank 0), and the
    elite's fitness never decreases."""
    with tempfile.TemporaryDirectory() as d:
        torch.manual_seed(0)
        torch.save(_G().state_dict(), os.path.join(d, 'G0.pt')); torch.save(_D().state_dict(), os.path.join(d, 'D0.pt'))
        _run_generation(0, 1, 0, d)
        mp.spawn(_run_generation, args=(2, _free_port(), d), nprocs=2, join=True)
        one = torch.load(os.path.join(d, 'gen_w1_r0.pt'))
        r0 = torch.load(os.path.join(d, 'gen_w2_r0.pt')); r1 = torch.load(os.path.join(d, 'gen_w2_r1.pt'))
    assert torch.equal(r0['final'], r1['final'])
    assert torch.allclose(r0['final'], one['final'], rtol=0, atol=0)
    for a, b, c in zip(one['history'], r0['history'], r1['history']):
        assert torch.equal(b, c) and torch.allclose(a, b, rtol=1e-6, atol=1e-7)
    best = [float(h.max()) for h in one['history']]
    assert best == sorted(best), best                                       # elitism: the best fitness is monotone
