"""Random-shape hunt of gg_upfirdn2d_f32 on the CPU build of the library (tests/emulated_lib.py) through the product's ctypes wrapper:
every kernel family behind the entry point (generic, tile, true-2x marching, unit-rate stream) as the dispatch picks them, against the
oracle.  Usage: python tools/fuzz_fir_on_model.py [cases] [seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import tests.util  # noqa: F401
from oracle import ops_ref as R
from tests import emulated_lib

plugin = emulated_lib.bind()
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
bad = done = 0
t_all = time.time()
for c in range(cases):
    four = rng.random() < 0.75
    f = R.setup_filter([1, 3, 3, 1]) if four else torch.from_numpy(rng.standard_normal((int(rng.integers(1, 6)), int(rng.integers(1, 6)))).astype(np.float32))
    up, down = ((1, 1), (2, 1), (1, 2), (2, 2), (3, 1), (1, 3))[int(rng.choice(6, p=[.3, .25, .25, .1, .05, .05]))]
    N, C = int(rng.integers(1, 3)), int(rng.integers(1, 5))
    H = int(rng.integers(1, 40))
    W = int(rng.choice([4, 8, 12, 16, 32, 64, 100, 128])) if rng.random() < 0.5 else int(rng.integers(1, 140))
    pad = [int(v) for v in rng.integers(-1, 4, size=4)]
    flip, gain = bool(rng.integers(0, 2)), float(rng.choice([1.0, 4.0, 0.5]))
    fh, fw = f.shape
    ow = (W * up + pad[0] + pad[1] - fw + down) // down
    oh = (H * up + pad[2] + pad[3] - fh + down) // down
    if ow < 1 or oh < 1:
        continue
    x = torch.from_numpy(rng.standard_normal((N, C, H, W)).astype(np.float32))
    want = R.upfirdn2d(x.double(), f, up=up, down=down, padding=pad, flip_filter=flip, gain=gain)
    y = plugin.upfirdn2d(x, f, up, up, down, down, pad[0], pad[1], pad[2], pad[3], flip, gain)
    err = float((y.double() - want).abs().max() / want.abs().max().clamp_min(1e-30))
    ok = tuple(y.shape) == tuple(want.shape) and err <= 5e-6
    done += 1
    if not ok or c % 10 == 0:
        print(f'case {c:3d} N{N} C{C} {H}x{W} f{fh}x{fw} up{up} down{down} pad{pad} flip{int(flip)} gain{gain}: {err:.1e}' + ('' if ok else '   <-- MISMATCH'), flush=True)
    bad += 0 if ok else 1
print(f'{done} cases run, {bad} mismatches, {time.time() - t_all:.0f}s')
sys.exit(1 if bad else 0)
