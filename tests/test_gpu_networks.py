"""GPU: network-level parity of the DROP-IN.  The reference's OWN Generator / Discriminator / StyleGAN2Loss (the checkout in
baseline/_ref, unmodified) run on this build's operators (gagan_b200.install) with the golden weights and must reproduce
what the same classes produced on the reference's CPU impl='ref' operators (tests/golden/networks.npz): eval and train
forward, and parameter gradients after each of the four loss phases (Greg and Dreg are double-backward through every op).
Every test runs twice: with the reference's forwards untouched, and with this build's fused forwards laid over them."""
import numpy as np
import pytest
import torch

from tests.util import load_golden, t, assert_close, patched_randn, TOL, reference_networks, quiet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', params=['fused_forwards', 'reference_forwards'])
def nets(device, request):
    import gagan_b200.training.networks as mine
    networks = reference_networks()
    mine.attach(networks, fused_callers=(request.param == 'fused_forwards'))
    g = load_golden('networks')
    cfg = {kv.split('=')[0]: int(kv.split('=')[1]) for kv in (str(m) for m in g['meta'])}
    G = networks.Generator(z_dim=cfg['z_dim'], c_dim=0, w_dim=cfg['w_dim'], img_resolution=cfg['res'], img_channels=3,
                           mapping_kwargs=dict(num_layers=cfg['num_layers']),
                           synthesis_kwargs=dict(channel_base=cfg['channel_base'], channel_max=cfg['channel_max']))
    D = networks.Discriminator(c_dim=0, img_resolution=cfg['res'], img_channels=3, channel_base=cfg['channel_base'],
                               channel_max=cfg['channel_max'], epilogue_kwargs=dict(mbstd_group_size=cfg['mbstd']))
    G.load_state_dict({k[2:]: t(v) for k, v in g.items() if k.startswith('G.')}, strict=False)
    D.load_state_dict({k[2:]: t(v) for k, v in g.items() if k.startswith('D.')}, strict=False)
    yield G.to(device), D.to(device), g, cfg
    mine.attach(networks, fused_callers=True)


def test_eval_forward(nets, device):
    G, D, g, cfg = nets
    G.eval(); D.eval()
    z = t(g['z'], device); c = torch.zeros(z.shape[0], 0, device=device)
    with torch.no_grad():
        ws = G.mapping(z, c)
        assert_close(ws, g['eval.ws'], TOL, 'mapping')
        img = G.synthesis(ws, noise_mode='const')
        assert_close(img, g['eval.img'], TOL, 'synthesis')
        assert_close(G(z, c, truncation_psi=0.7, truncation_cutoff=4, noise_mode='const'), g['eval.img_trunc'], TOL, 'trunc')
        assert_close(D(t(g['eval.img'], device), c), g['eval.logits'], TOL, 'D')


def test_train_forward_random_noise(nets, device):
    G, D, g, cfg = nets
    G.train()
    G.mapping.w_avg_beta = None
    with torch.no_grad(), patched_randn(7):
        img = G.synthesis(t(g['eval.ws'], device), noise_mode='random')
    assert_close(img, g['train.img_randnoise7'], TOL)


@pytest.mark.parametrize('phase', ['Gmain', 'Greg', 'Dmain', 'Dreg'])
def test_loss_phase_parameter_gradients(nets, device, phase):
    from training.loss import StyleGAN2Loss
    G, D, g, cfg = nets
    G.train(); D.train()
    G.mapping.w_avg_beta = None
    for p in list(G.parameters()) + list(D.parameters()):
        p.requires_grad_(True)
        p.grad = None
    loss = StyleGAN2Loss(device=device, G_mapping=G.mapping, G_synthesis=G.synthesis, D=D, style_mixing_prob=0,
                         r1_gamma=10, pl_batch_shrink=2, pl_decay=0.01, pl_weight=2)
    z = t(g['z'], device); real = t(g['real'], device); c = torch.zeros(z.shape[0], 0, device=device)
    with patched_randn(11):
        loss.accumulate_gradients(phase=phase, real_img=real, real_c=c, gen_z=z, gen_c=c, sync=True, gain=1.0)
    net = G if phase[0] == 'G' else D
    worst = ('', 0.0)
    for k, p in net.named_parameters():
        want = g[f'{phase}.grad.{k}']
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        if np.abs(want).max() == 0:
            assert float(got.abs().max()) <= 1e-6, k
            continue
        e = assert_close(got, want, TOL, f'{phase}.{k}')
        if e > worst[1]:
            worst = (k, e)
    if phase == 'Greg':
        assert_close(loss.pl_mean, g['Greg.pl_mean'], TOL, 'pl_mean')
    print(f'{phase}: worst max-rel-err {worst[1]:.2e} at {worst[0]}')


def test_ga_population_fitness_eval_on_the_device(device):
    """BASELINE configs[3] at toy size: fitness of StyleSpace-offset individuals through the CUDA G/D (world = 1)."""
    from gagan_b200.training import ga_eval
    networks = reference_networks()
    torch.manual_seed(3)
    G = quiet(networks.Generator, z_dim=32, c_dim=0, w_dim=32, img_resolution=32, img_channels=3, mapping_kwargs=dict(num_layers=2),
                           synthesis_kwargs=dict(channel_base=512, channel_max=32, use_domain_modulation=True,
                                                 domain_modulation_parametrization='additive')).to(device)
    D = networks.Discriminator(c_dim=0, img_resolution=32, img_channels=3, channel_base=512, channel_max=32).to(device)
    n_layers = len(ga_eval.offset_layers(G))
    assert n_layers == 4 * 2 + 4 - 1                                       # conv0/conv1/torgb per block, no conv0 in b4
    assert ga_eval.genome_size(G) == sum(m.weight.shape[1] for _, m in ga_eval.offset_layers(G))   # one offset per input channel
    pop = ga_eval.init_population(G, size=5, scale=0.1, seed=1)
    pop[2].zero_()                                                          # individual 2 = the unmodified generator
    z = torch.randn(4, 32, device=device)
    fit = ga_eval.evaluate_population(G, D, pop, z)
    assert fit.shape == (5,) and torch.isfinite(fit).all()
    G.eval(); D.eval()
    with torch.no_grad():
        ga_eval.load_individual(G, torch.zeros(ga_eval.genome_size(G)))
        c = torch.zeros(4, 0, device=device)
        base = D(G.synthesis(G.mapping(z, c), noise_mode='const'), c).mean()
    assert abs(float(fit[2] - base)) <= 1e-5 * max(1.0, abs(float(base)))
    assert len({round(float(v), 5) for v in fit}) == 5                      # different offsets, different fitness


def test_ga_population_eval_cuda_graph_replay_matches_eager(device):
    """Shards of >= 16 individuals are evaluated by replaying one captured CUDA graph (ga_eval.evaluate_population): the
    same kernels, so the fitness vector must equal the eager loop's bit for bit (up to atomics-free determinism)."""
    from gagan_b200.training import ga_eval
    networks = reference_networks()
    torch.manual_seed(4)
    G = quiet(networks.Generator, z_dim=32, c_dim=0, w_dim=32, img_resolution=32, img_channels=3, mapping_kwargs=dict(num_layers=2),
                           synthesis_kwargs=dict(channel_base=512, channel_max=32, use_domain_modulation=True,
                                                 domain_modulation_parametrization='additive')).to(device)
    D = networks.Discriminator(c_dim=0, img_resolution=32, img_channels=3, channel_base=512, channel_max=32).to(device)
    pop = ga_eval.init_population(G, size=18, scale=0.1, seed=2)
    z = torch.randn(4, 32, device=device)
    eager = ga_eval.evaluate_population(G, D, pop, z, cuda_graph=False)
    graph = ga_eval.evaluate_population(G, D, pop, z, cuda_graph=True)
    again = ga_eval.evaluate_population(G, D, pop, z, cuda_graph=True)       # a second capture in the same process
    assert torch.isfinite(graph).all() and len({round(float(v), 5) for v in graph}) == 18
    assert float((graph - eager).abs().max()) <= 1e-6 * float(eager.abs().max())
    assert float((again - graph).abs().max()) <= 1e-6 * float(eager.abs().max())
