#!/usr/bin/env python
"""BASELINE configs[0] on the device: 256^2 paper256 random-init generator, synthesis forward, batch 4, eval, noise_mode='const'
(the reference's CPU-runnable case: 0.725 s per batch with impl='ref' on 8 cores)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gagan_b200  # noqa: E402
_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(_CHECKOUT if os.path.isdir(_CHECKOUT) else None)   # the reference checkout on this build's operators
import torch
from torch_utils import custom_ops
from gagan_b200.training import training_loop
custom_ops.verbosity = 'none'
dev = torch.device('cuda:0')
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
G, _ = training_loop.build_networks(256, 'paper256', device=dev)
G.eval()
for N in (4, 32):
    z = torch.randn(N, 512, device=dev); c = torch.zeros(N, 0, device=dev)
    with torch.no_grad():
        for _ in range(3):
            img = G(z, c, noise_mode='const')
        torch.cuda.synchronize()
        l0 = custom_ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            img = G(z, c, noise_mode='const')
        e1.record(); torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) / 10
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            img_g = G(z, c, noise_mode='const')
        g.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        graph = e0.elapsed_time(e1) / 10
    print(f'G 256^2 paper256 forward, batch {N}: eager {eager:.2f} ms ({N / eager * 1e3:.0f} img/s, {(custom_ops.launch_count() - l0) // 12} library launches), '
          f'CUDA graph replay {graph:.2f} ms ({N / graph * 1e3:.0f} img/s); image {tuple(img.shape)} finite={bool(torch.isfinite(img).all())} '
          f'graph==eager: {bool(torch.equal(img, img_g))}', flush=True)
