"""Random-shape hunt on the CPU build of the library (tests/emulated_lib.py): gg_conv2d_f32 / gg_conv2d_wgrad_f32 through the product's
own ctypes wrappers, AUTO dispatch (tcgen05 tile / marching / thin / FFMA kernels, on the hardware model), against float64 torch.
Usage: python tools/fuzz_on_model.py [cases] [seed]   -- prints one line per case and a summary; exit code 1 on any mismatch."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F
import tests.util  # noqa: F401  (installs the drop-in)
from tests import emulated_lib

plugin = emulated_lib.bind()
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bad = 0
t_all = time.time()
for c in range(cases):
    K = int(rng.choice([1, 2, 3]))
    N = int(rng.integers(1, 4))
    I = int(rng.choice([3, 8, 16, 17, 24, 32, 40, 64]))
    O = int(rng.choice([3, 16, 20, 32, 33, 64, 72, 130]))
    wide = rng.random() < 0.3
    H = int(rng.integers(K, 30))
    W = int(rng.choice([64, 68, 72, 132])) if wide else int(rng.integers(max(K, 2), 40))
    py, px = int(rng.integers(0, K)), int(rng.integers(0, K))
    flip, transposed, scales = bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    g = torch.Generator().manual_seed(c)
    x = torch.randn(N, I, H, W, generator=g)
    w = torch.randn((I, O, K, K) if transposed else (O, I, K, K), generator=g)
    si = torch.randn(N, I, generator=g) if scales else None
    so = torch.randn(N, O, generator=g) if scales else None
    xd = x.double() * (si.double()[:, :, None, None] if scales else 1)
    wd = w.double().flip([2, 3]) if flip else w.double()
    if transposed:
        if py > K - 1 or px > K - 1:
            continue
        want = F.conv_transpose2d(xd, wd, padding=(py, px))
    else:
        if H + 2 * py < K or W + 2 * px < K:
            continue
        want = F.conv2d(xd, wd, padding=(py, px))
    if scales:
        want = want * so.double()[:, :, None, None]
    t0 = time.time()
    y = plugin.conv2d(x, w, stride=1, padding=(py, px), transposed=transposed, flip_w=flip, in_scale=si, out_scale=so)
    prec = plugin.last_conv_prec
    err = float((y.double() - want).abs().max() / want.abs().max().clamp_min(1e-30))
    # weight gradient of the correlation y = conv(a, w): a = x, b = random gradient of the natural output
    line = f'case {c:3d} N{N} I{I} O{O} {H}x{W} k{K} pad({py},{px}) flip{int(flip)} T{int(transposed)} s{int(scales)} prec {prec}: conv {err:.1e}'
    ok = err <= 1e-5
    if not transposed:
        b = torch.randn(want.shape, generator=g)
        wv = torch.zeros(O, I, K, K, dtype=torch.float64, requires_grad=True)
        bd = b.double() * (so.double()[:, :, None, None] if scales else 1)
        (F.conv2d(xd, wv, padding=(py, px)) * bd).sum().backward()
        gw = wv.grad.flip([2, 3]) if flip else wv.grad
        dw = plugin.conv2d_wgrad(x, b, (K, K), stride=1, padding=(py, px), flip_w=flip, out_layout=0, a_scale=si, b_scale=so)
        scale = float(((xd ** 2).sum() * (bd ** 2).sum() / (I * O)).sqrt())
        werr = float((dw.double() - gw).abs().max()) / max(scale, float(gw.abs().max()))
        line += f'  wgrad(prec {plugin.last_wgrad_prec}) {werr:.1e}'
        ok = ok and werr <= 1e-5
    print(line + f'  {time.time() - t0:.1f}s' + ('' if ok else '   <-- MISMATCH'), flush=True)
    bad += 0 if ok else 1
print(f'{cases} cases, {bad} mismatches, {time.time() - t_all:.0f}s')
sys.exit(1 if bad else 0)
