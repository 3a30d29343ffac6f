"""A thread-level model of fir_stream<0, PX0, STG> (ga-gan_b200/csrc/upfirdn2d.cu): the launch arithmetic (column groups, padding to whole
warps, strips), the coalesced input window with its shared-memory exchange, and the coalesced output exchange through per-lane row
pointers -- executed lane by lane in numpy and compared with the oracle's upfirdn2d.  It pins the INDEX ARITHMETIC of the kernel for
shapes whose rows span several warps (the GPU tests of the same kernel are tests/test_gpu_ops.py::test_upfirdn2d_*)."""
import numpy as np
import pytest
import torch

from oracle import ops_ref as R


def _model(x, f, padx0, pady0, flip, gain, outH, outW, force_staged=False):
    N, C, inH, inW = x.shape
    vec_in, vec_out = inW % 4 == 0, outW % 4 == 0
    fullW, fullH = outW, outH
    RS = 32 if fullH >= 128 else (16 if fullH >= 32 else 8)
    ncg, nst = (fullW + 7) // 8, (fullH + RS - 1) // RS
    staged = force_staged or (not vec_in and fullW >= 256) or (not vec_out and fullW >= 384)
    assert staged, 'the model covers the staged kernel'
    if not vec_in:
        ncg = (ncg + 31) // 32 * 32                                  # launch_stream2: whole warps per row for the input window
    threads = ncg * nst * N * C
    blocks = (threads + 127) // 128
    K = np.zeros((4, 4), np.float64)
    for ky in range(4):
        for kx in range(4):
            K[ky, kx] = gain * f[ky if flip else 3 - ky, kx if flip else 3 - kx]
    PX0 = padx0
    raw_rows = not vec_in
    out = np.full((N, C, outH, outW), np.nan)
    park = (max(inW, outW) + 31) // 8 * 8 + 64
    lane = np.arange(32)
    for warp in range(blocks * 4):
        ids = warp * 32 + lane
        cg = ids % ncg; r = ids // ncg; st = r % nst; nc = r // nst
        past = nc >= N * C
        n = np.where(past, 0, nc // C); c = np.where(past, 0, nc % C)
        x0 = np.where(past, park, cg * 8); y0 = st * RS
        nt = RS + 3
        rows = []                                                     # in[lane][11] of the last four steps
        for rr in range(nt):
            rin = y0 - pady0 + rr
            row_ok = (rin >= 0) & (rin < inH)
            inn = np.zeros((32, 11))
            if raw_rows:
                sbuf = np.zeros(288)
                w0 = x0 - 8 * lane - 4
                for i in range(9):
                    col = w0 + 32 * i + lane
                    ok = row_ok & (col >= 0) & (col < inW)
                    val = np.where(ok, x[n, c, np.where(row_ok, rin, 0), np.clip(col, 0, inW - 1)], 0.0)
                    sbuf[32 * i + lane] = val
                for j in range(11):
                    inn[:, j] = sbuf[8 * lane + 4 - PX0 + j]
            else:
                for j in range(11):
                    col = x0 - PX0 + j
                    ok = row_ok & (col >= 0) & (col < inW)
                    inn[:, j] = np.where(ok, x[n, c, np.where(row_ok, rin, 0), np.clip(col, 0, inW - 1)], 0.0)
            rows.append(inn)
            if rr < 3:
                continue
            y = y0 + rr - 3
            o = np.zeros((32, 8))
            for ky in range(4):
                src = rows[rr - 3 + ky]
                for kx in range(4):
                    o += K[ky, kx] * src[:, kx:kx + 8]
            if not vec_out:                                           # staged stores: per-lane row pointer + valid count, coalesced write-out
                v = o.reshape(-1).copy()                              # sb.v[8 lane + t]
                nv = np.where(y < outH, np.clip(outW - x0, 0, 8), 0)
                for t in range(8):
                    e = 32 * t + lane; srcl = e >> 3; k = e & 7
                    for l_ in range(32):
                        s_ = srcl[l_]
                        if k[l_] < nv[s_]:
                            out[n[s_], c[s_], y[s_], x0[s_] + k[l_]] = v[e[l_]]
            else:
                for l_ in range(32):
                    if y[l_] < outH:
                        for t in range(8):
                            if x0[l_] + t < outW:
                                out[n[l_], c[l_], y[l_], x0[l_] + t] = o[l_, t]
    return out


@pytest.mark.parametrize('case', [
    # N, C, inH, inW, padding [x0, x1, y0, y1], flip, gain
    (1, 2, 20, 515, [1, 1, 1, 1], False, 4.0),        # both sides unaligned, 3 warps per row (65 column groups padded to 96)
    (1, 1, 37, 257, [1, 1, 1, 1], True, 1.0),         # odd input -> 256 columns: exactly one warp per row, two strips
    (2, 1, 9, 384, [2, 2, 2, 2], False, 1.0),         # aligned input -> 385 columns: output exchange only, warps span rows and planes
    (1, 1, 12, 1025, [1, 1, 2, 0], False, 2.0),       # 128 column groups = 4 warps per row
    (1, 3, 10, 131, [3, 0, 0, 3], True, 1.0),         # narrow (forced): padx0 = 3, 17 groups padded to 32
    (1, 1, 40, 300, [0, 3, 3, 0], False, 1.0),        # padx0 = 0; aligned input, unaligned output (forced)
])
def test_staged_fir_index_arithmetic(case):
    N, C, H, W, pad, flip, gain = case
    g = torch.Generator().manual_seed(H * 1000 + W)
    x = torch.randn(N, C, H, W, generator=g, dtype=torch.float64)
    f = R.setup_filter([1, 3, 3, 1])                       # float32 coefficients (exact binary fractions), as the product passes them
    want = R.upfirdn2d(x, f, padding=pad, flip_filter=flip, gain=gain).numpy()
    outH, outW = want.shape[2:]
    got = _model(x.numpy(), f.double().numpy(), pad[0], pad[2], flip, gain, outH, outW, force_staged=True)
    assert not np.isnan(got).any(), 'some output element was never written'
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
