"""The G+D training iteration the images/sec metric is defined over, driving the REFERENCE's networks and loss.

The reference's `training_loop()` (DissimilarDomains/training/training_loop.py:163-666) is a 500-line driver whose
optimiser step / EMA code was de-indented out of its loop in this fork; the intended behaviour
is upstream stylegan2-ada-pytorch.  `TrainingStep` is that inner iteration and nothing else:

    phases Gmain, Greg (every G_reg_interval), Dmain, Dreg (every D_reg_interval) with the lazy-regularisation
    rescaling of lr and betas                                                                  (:293-318)
    per phase: zero_grad, requires_grad on the phase's trainable PARTS only (`select_parts`, the reference's
    name_filters / set_requires_grad :57-95), `batch_size // (batch_gpu*num_gpus)` accumulation rounds of the
    reference's own `StyleGAN2Loss.accumulate_gradients`, nan_to_num on grads, Adam step       (:459-512)
    G_ema <- lerp(G, G_ema, 0.5 ** (batch_size / ema_nimg)), buffers copied                    (:515-523)
    ADA heuristic on `Loss/signs/real` every `ada_interval` iterations                         (:526-533)

Networks, loss and the ADA pipe are the reference checkout's classes (training/networks.py, training/loss.py,
training/augment.py) running on this build's operators -- `gagan_b200.install(checkout)` must have been called.
Data loading, snapshots, metrics, logging and the GA hooks are out of scope.  With torch.distributed initialised the
modules are wrapped in DistributedDataParallel where :270-285 wraps them (one process per GPU, NCCL all-reduce of the
phase's gradients on its last round).
"""
import io
import copy
import importlib
import contextlib
import numpy as np
import torch

CONFIGS = {
    # train.py:219-228 (+ --fp32: num_fp16_res=0, conv_clamp=None, :418-423)
    'stylegan2': dict(fmaps=1.0, lrate=0.002, gamma=10.0, ema=10, mbstd=4, map=8, mb=32),   # config-f
    'paper256':  dict(fmaps=0.5, lrate=0.0025, gamma=1.0, ema=20, mbstd=8, map=8, mb=64),
    'paper512':  dict(fmaps=1.0, lrate=0.0025, gamma=0.5, ema=20, mbstd=8, map=8, mb=64),
    'paper1024': dict(fmaps=1.0, lrate=0.002, gamma=2.0, ema=10, mbstd=4, map=8, mb=32),
}

AUGPIPE_BGC = dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1, brightness=1, contrast=1,
                   lumaflip=1, hue=1, saturation=1)        # train.py:365-368 ('bgc', the ADA default)


def _ref(module):
    """A module of the installed reference checkout (`training.networks`, `training.loss`, `torch_utils.misc`, ...)."""
    import gagan_b200
    if gagan_b200.installed()['reference_root'] is None:
        raise RuntimeError('gagan_b200.install(<DissimilarDomains checkout>) must run first: the training step drives the '
                           "reference's own networks and loss on this build's operators")
    return importlib.import_module(module)


def build_networks(resolution, cfg='stylegan2', z_dim=512, w_dim=512, channel_max=512, device='cuda', **synthesis_extra):
    """G and D as train.py:264-273 builds them for `--cfg=<cfg> --fp32=1`; `synthesis_extra` carries the domain-adaptation
    options (use_domain_modulation, domain_modulation_parametrization, generator_requires_grad_parts; train.py:459-463)."""
    networks = _ref('training.networks')
    spec = CONFIGS[cfg]
    channel_base = int(spec['fmaps'] * 32768)
    with contextlib.redirect_stdout(io.StringIO()):          # register_*_modulation print one line per layer
        G = networks.Generator(z_dim=z_dim, c_dim=0, w_dim=w_dim, img_resolution=resolution, img_channels=3,
                               mapping_kwargs=dict(num_layers=spec['map']),
                               synthesis_kwargs=dict(channel_base=channel_base, channel_max=channel_max, num_fp16_res=0,
                                                     conv_clamp=None, **synthesis_extra))
        D = networks.Discriminator(c_dim=0, img_resolution=resolution, img_channels=3, channel_base=channel_base,
                                   channel_max=channel_max, num_fp16_res=0, conv_clamp=None, block_kwargs={}, mapping_kwargs={},
                                   epilogue_kwargs=dict(mbstd_group_size=spec['mbstd']))
    return G.to(device), D.to(device)


# ----------------------------------------------------------------------------
# Trainable parts of the generator (training_loop.py:57-95).  A part is `<group>` or `<group>.b<res>`; every group is a
# conjunction of substring tests on the parameter name, written here as (scope, all-of, any-of, none-of).

_PART_RULES = {
    'mapping':                    ('', ('mapping',), (), ()),
    'tRGB_affine':                ('synthesis', ('torgb.affine',), (), ()),
    'tRGB_conv':                  ('synthesis', (), ('torgb.weight', 'torgb.bias'), ('affine', 'offset')),
    'tRGB_offset':                ('synthesis', ('torgb.offset',), (), ('torgb.weights_offset',)),
    'tRGB_weights_offset':        ('synthesis', ('torgb.weights_offset',), (), ()),
    'tRGB_affine_weights_offset': ('synthesis', ('torgb.affine.weights_offset',), (), ()),
    'synt_affine':                ('synthesis', ('conv', 'affine'), (), ()),
    'synt_conv':                  ('synthesis', ('conv',), ('weight', 'noise_strength', 'bias'), ('affine', 'offset')),
    'synt_const':                 ('synthesis', ('const',), (), ()),
    'synt_offset':                ('synthesis', ('conv', 'offset'), (), ('weights_offset',)),
    'synt_weights_offset':        ('synthesis', ('conv', 'weights_offset'), (), ('affine',)),
    'synt_affine_weights_offset': ('synthesis', ('conv', 'affine.weights_offset'), (), ()),
}


_PART_RESOLUTIONS = (1024, 512, 256, 128, 64, 32, 16, 8, 4)


def select_parts(module, parts):
    """Names of the parameters of `module` that `parts` makes trainable ('all' = every parameter)."""
    names = [n for n, _ in module.named_parameters()]
    if 'all' in parts:
        return names
    chosen = set()
    for part in parts:
        group, _, res = part.partition('.b')
        if res and not res.isdigit():
            continue
        if group not in _PART_RULES:
            continue                                        # the reference silently ignores unknown part names
        scope, all_of, any_of, none_of = _PART_RULES[group]
        if res and int(res) not in _PART_RESOLUTIONS:
            continue
        if group != 'mapping':                              # the reference's 'mapping' filter ignores the block suffix
            scope = f'synthesis.b{res}' if res else scope
        for n in names:
            if scope in n and all(s in n for s in all_of) and (not any_of or any(s in n for s in any_of)) \
                    and not any(s in n for s in none_of):
                chosen.add(n)
    return [n for n in names if n in chosen]


def set_requires_grad(module, parts):
    wanted = set(select_parts(module, parts))
    for n, p in module.named_parameters():
        p.requires_grad_(n in wanted)


# ----------------------------------------------------------------------------

class TrainingStep:
    def __init__(self, G, D, batch_size, batch_gpu, device, lrate=0.002, glrate=None, r1_gamma=10.0, ema_kimg=10.0, G_reg_interval=4,
                 D_reg_interval=16, style_mixing_prob=0.9, pl_weight=2.0, rank=0, num_gpus=1, g_parts=('all',),
                 augment_kwargs=None, augment_p=0.0, ada_target=None, ada_interval=4, ada_kimg=500):
        assert batch_size % (batch_gpu * num_gpus) == 0
        misc = _ref('torch_utils.misc')
        self._misc = misc
        self._stats = _ref('torch_utils.training_stats')
        loss_mod = _ref('training.loss')
        conv2d_gradfix = _ref('torch_utils.ops.conv2d_gradfix')
        self.device = torch.device(device)
        self.batch_size, self.batch_gpu, self.num_gpus, self.rank = batch_size, batch_gpu, num_gpus, rank
        self.G = G.train().requires_grad_(False).to(self.device)
        self.D = D.train().requires_grad_(False).to(self.device)
        self.G_ema = copy.deepcopy(self.G).eval()
        self.ema_nimg = ema_kimg * 1000
        conv2d_gradfix.enabled = True                          # training_loop.py:209
        _ref('torch_utils.ops.grid_sample_gradfix').enabled = True   # :210
        torch.backends.cuda.matmul.allow_tf32 = False          # :207-208
        torch.backends.cudnn.allow_tf32 = False
        self.parts = dict(G=list(g_parts), D=['all'])          # train.py:451-458
        if num_gpus > 1 and not self._stats._sync_called:
            self._stats.init_multiprocessing(rank=rank, sync_device=self.device)

        # ADA (:247-258)
        self.augment_pipe, self.ada_stats = None, None
        self.ada_target, self.ada_interval, self.ada_kimg = ada_target, ada_interval, ada_kimg
        if augment_kwargs is not None and (augment_p > 0 or ada_target is not None):
            augment = _ref('training.augment')
            self.augment_pipe = augment.AugmentPipe(**augment_kwargs).train().requires_grad_(False).to(self.device)
            self.augment_pipe.p.copy_(torch.as_tensor(augment_p))
            if ada_target is not None:
                self.ada_stats = self._stats.Collector(regex='Loss/signs/real')

        # DDP wrap (:270-285).  G_ema is not wrapped.  Only what the phases train is handed to the reducer: with a part
        # filter the reference's "all parameters" wrap would leave buckets waiting for gradients that never come.
        ddp = dict(G_mapping=self.G.mapping, G_synthesis=self.G.synthesis, D=self.D)
        if num_gpus > 1:
            set_requires_grad(self.G, self.parts['G'])
            self.D.requires_grad_(True)
            for name, module in list(ddp.items()):
                if any(p.requires_grad for p in module.parameters()):
                    module = torch.nn.parallel.DistributedDataParallel(
                        module, device_ids=[self.device] if self.device.type == 'cuda' else None, broadcast_buffers=False)
                ddp[name] = module
            self.G.requires_grad_(False)
            self.D.requires_grad_(False)
        self.loss = loss_mod.StyleGAN2Loss(device=self.device, **ddp, augment_pipe=self.augment_pipe,
                                           style_mixing_prob=style_mixing_prob, r1_gamma=r1_gamma, pl_weight=pl_weight)

        # Phases with lazy regularisation (:293-318); the optimiser holds every parameter, Adam skips those without a gradient.
        self.phases = []
        for name, module, lr, reg_interval in [('G', self.G, glrate if glrate is not None else lrate, G_reg_interval),
                                               ('D', self.D, lrate, D_reg_interval)]:
            params = list(module.parameters())
            if reg_interval is None:
                opt = torch.optim.Adam(params, lr=lr, betas=(0.0, 0.99), eps=1e-8)
                self.phases.append(dict(name=name + 'both', net=name, module=module, opt=opt, interval=1))
            else:
                mb_ratio = reg_interval / (reg_interval + 1)
                opt = torch.optim.Adam(params, lr=lr * mb_ratio, betas=(float(0 ** mb_ratio), 0.99 ** mb_ratio), eps=1e-8)
                self.phases.append(dict(name=name + 'main', net=name, module=module, opt=opt, interval=1))
                self.phases.append(dict(name=name + 'reg', net=name, module=module, opt=opt, interval=reg_interval))
        self.loss_stats = self._stats.Collector(regex='Loss/.*')
        self.cur_it = 0

    def run(self, real_img, gen_z_all=None):
        """One iteration.  real_img: [batch_size // num_gpus, 3, R, R] float32 in [-1, 1] on the device.
        gen_z_all: optional [len(phases), batch_size // num_gpus, z_dim]; drawn on the device when None (:446-451)."""
        per_gpu = self.batch_size // self.num_gpus
        assert real_img.shape[0] == per_gpu
        real_c = torch.zeros([per_gpu, 0], device=self.device)
        if gen_z_all is None:
            gen_z_all = torch.randn([len(self.phases), per_gpu, self.G.z_dim], device=self.device)
        real_rounds = real_img.split(self.batch_gpu)
        c_rounds = real_c.split(self.batch_gpu)
        for phase, phase_z in zip(self.phases, gen_z_all):
            if self.cur_it % phase['interval'] != 0:
                continue
            phase['opt'].zero_grad(set_to_none=True)
            set_requires_grad(phase['module'], self.parts[phase['net']])
            z_rounds = phase_z.split(self.batch_gpu)
            for round_idx, (r_img, r_c, g_z) in enumerate(zip(real_rounds, c_rounds, z_rounds)):
                sync = (round_idx == len(real_rounds) - 1)
                self.loss.accumulate_gradients(phase=phase['name'], real_img=r_img, real_c=r_c, gen_z=g_z, gen_c=r_c,
                                               sync=sync, gain=phase['interval'])
            phase['module'].requires_grad_(False)
            grads = [p.grad for p in phase['module'].parameters() if p.grad is not None]
            for g in grads:
                torch.nan_to_num(g, nan=0, posinf=1e5, neginf=-1e5, out=g)
            phase['opt'].step()

        # G_ema (:515-523)
        ema_beta = 0.5 ** (self.batch_size / max(self.ema_nimg, 1e-8))
        with torch.no_grad():
            p_ema, p_cur = list(self.G_ema.parameters()), list(self.G.parameters())
            torch._foreach_lerp_(p_ema, p_cur, 1.0 - ema_beta)       # p_ema = lerp(p, p_ema, beta)
            for b_ema, b in zip(self.G_ema.buffers(), self.G.buffers()):
                b_ema.copy_(b)
        self.cur_it += 1

        # ADA heuristic (:526-533)
        if self.ada_stats is not None and self.cur_it % self.ada_interval == 0:
            self.ada_stats.update()
            adjust = np.sign(self.ada_stats['Loss/signs/real'] - self.ada_target) * (self.batch_size * self.ada_interval) / (self.ada_kimg * 1000)
            self.augment_pipe.p.copy_((self.augment_pipe.p + adjust).max(self._misc.constant(0, device=self.device)))

    def read_stats(self):
        """The step's loss statistics on the HOST (device -> host copy, and the cross-rank all-reduce of training_stats in a
        multi-GPU run): what the reference's `stats_collector.update()` does once per tick (:545-560)."""
        self.loss_stats.update()
        return self.loss_stats.as_dict()
