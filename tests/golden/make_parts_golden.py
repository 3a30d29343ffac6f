#!/usr/bin/env python
"""Golden selections of the reference's trainable-part filter (training_loop.py:57-95), produced by the reference itself.

Needs /root/reference (the function lives in training/training_loop.py, which imports the checkout's `legacy` and `metrics`
top-level modules that are not vendored).  Writes tests/golden/requires_grad_parts.json.
"""
import io
import os
import sys
import json
import contextlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference/DissimilarDomains')

from training import networks                                   # noqa: E402  (reference)
from training.training_loop import set_requires_grad            # noqa: E402  (reference)

PART_SPECS = [['all'], ['synt_affine', 'tRGB_affine', 'synt_weights_offset.b64', 'tRGB_weights_offset.b64'],      # README.md:191-196 (Affine+)
              ['synt_offset', 'tRGB_offset'], ['mapping'], ['mapping.b64'], ['synt_conv.b8', 'tRGB_conv'], ['synt_const'],
              ['synt_affine_weights_offset', 'tRGB_affine_weights_offset.b16'], ['synt_weights_offset', 'nonsense', 'synt_conv.b3'], [],
              ['tRGB_offset.b32', 'synt_affine.b4', 'synt_weights_offset.b2048']]

with contextlib.redirect_stdout(io.StringIO()):
    G = networks.Generator(z_dim=16, c_dim=0, w_dim=16, img_resolution=64, img_channels=3, mapping_kwargs=dict(num_layers=2),
                           synthesis_kwargs=dict(channel_base=1024, channel_max=16, use_domain_modulation=True,
                                                 domain_modulation_parametrization='out_in_5_1,additive,affine_out_in_5_1',
                                                 generator_requires_grad_parts=['all']))
selected = []
for spec in PART_SPECS:
    set_requires_grad(G, list(spec))
    selected.append([n for n, p in G.named_parameters() if p.requires_grad])
out = dict(parameter_names=[n for n, _ in G.named_parameters()], specs=[list(s) for s in PART_SPECS], selected=selected)
json.dump(out, open(os.path.join(HERE, 'requires_grad_parts.json'), 'w'), indent=0)
print('wrote requires_grad_parts.json:', [len(s) for s in selected])
