// tcgen05 / TMEM weight-gradient kernel for sm_100a (stride-1 convolutions, k in {1,2,3}), fp32-faithful 3xTF32.
//
// Replaces aten::cudnn_convolution_backward_weight behind conv2d_gradfix.Conv2dGradWeight (torch_utils/ops/
// conv2d_gradfix.py:175-191) for every stride-1 layer and, through the phase-major formulation of conv2d_resample.py,
// for the stride-2 layers as well:
//
//   dw[b,a,ky,kx] = sum_{n,y,x} (gs[n,b] * G[n,b,y,x]) * (xs[n,a] * X[n,a,y+ky-py,x+kx-px])          X == 0 outside its extent
//
// GEMM view     D_ky[m = (kx slot, grad channel b), n = input channel a] += sum_{pixels} Gshift[m, pix] * X_ky[n, pix]
//               The reduction dimension is the PIXEL axis, which is the contiguous axis of both NCHW operands, so both
//               are K-major UMMA operands: shared-memory image [4-pixel chunk][row][4 pixels] (no-swizzle canonical layout).
//               A ky shift of X is a different row of the X ring (= a different descriptor start address); a kx shift
//               cannot be expressed by a descriptor (it breaks the 16-byte chunk alignment), so the G operand is stored
//               once per kx with the shift applied while it is converted -- with few channels several kx copies are
//               stacked into the 128 rows of one MMA (32 channels: 3 copies, 64: 2, >= 128: one kx per CTA group).
// Work unit     one (tile, strip): tile = (128 rows of (kx,b)) x (NTA input channels) x (all ky); strip = RR image rows x
//               16 image columns of one sample.  A persistent CTA walks a contiguous range of units.
// Pipeline      converter warps load G / X rows straight from global memory (coalesced 64-byte runs, 128-bit loads when
//               aligned), apply the per-sample scales, split into tf32 hi/lo (cvt.rna) and fill two shared-memory rings;
//               one thread issues tcgen05.mma.kind::tf32 (M=128, N=NTA, K=8; hi*hi + hi*lo + lo*hi) into TMEM.
// Accumulation  as in conv_tc.cu the tensor core's truncating fp32 accumulator is only trusted for one strip
//               (<= 16 rows x 2 steps x 3 products); two TMEM sets ping-pong and the converter warps drain every finished
//               strip into fp32 REGISTERS (round-to-nearest).  A tile's partial sum leaves the CTA once, with fp32 atomics
//               into the zero-initialised dw (a handful of partials per tile: one per CTA that worked on it).
#define GG_TU_TAG 3
#include "tc_common.cuh"
#include <stdlib.h>

using namespace ggtc;

namespace {

constexpr int WG_CONS_WARPS = 8;
constexpr int WG_CONS_THREADS = WG_CONS_WARPS * 32;
constexpr int WG_GROUP_WARPS = 4;                        // the converter warps work as two groups that alternate tasks
constexpr int WG_GROUP_THREADS = WG_GROUP_WARPS * 32;
constexpr int WG_PROD_WARPS = 4;                         // one warpgroup (w0 = MMA issuer + TMEM owner, w1..w3 idle) so that setmaxnreg applies
constexpr int WG_THREADS = (WG_PROD_WARPS + WG_CONS_WARPS) * 32;   // w4..w11 = convert + drain + flush
constexpr int UW = 16;                                   // image columns per strip (two K=8 steps)
constexpr int GS = 4;                                    // G-row ring slots (power of two: slot = counter & (GS-1))
constexpr int XS = 8;                                    // X-row ring slots (k live rows + rows in flight; power of two)
constexpr uint32_t LBO_A = 128 * 16;                     // chunk pitch of the G image: every 8-row core matrix stays 128-byte aligned
// Item -> (row, chunk) mapping of the converter threads: 8 consecutive lanes write 8 consecutive rows of ONE chunk (a 128-byte
// conflict-free quarter-warp store), the four quarter-warps take the four chunks; in global memory the same warp reads 8
// channels x 64 contiguous bytes.
__device__ __forceinline__ int item_chunk(int id) { return (id >> 3) & 3; }
__device__ __forceinline__ int item_row(int id) { return (id & 7) | ((id >> 5) << 3); }

struct WgP {
    const float* X; const float* G; float* dw; const float* xs; const float* gs;
    int N, A, HA, WA, B, HB, WB, K, pad_y, pad_x, flip_w, out_layout;
    int RB, rb_shift, nshift, zgroups, btiles, atiles;
    int RR, ustrips, rstrips, S;
    int total_units, units_per_cta, nprod, vecX, vecG;
    int dbg;                    // always 0 in the library (the role-ablation bits of the development builds: 1 = no MMAs, 2 = no global loads, 4 = no split / st.shared, 8 = no proxy fence)
};

struct Unit { int n, r0, rows, u0, b0, a0, kx0, ns, tile; };

__device__ __forceinline__ Unit decode_unit(int unit, const WgP& p, int NTA) {
    Unit u;
    u.tile = unit / p.S;
    int strip = unit - u.tile * p.S;
    int us = strip % p.ustrips; strip /= p.ustrips;
    int rs = strip % p.rstrips;
    u.n = strip / p.rstrips;
    u.u0 = us * UW;
    u.r0 = rs * p.RR;
    u.rows = min(p.RR, p.HB - u.r0);
    int t = u.tile;
    int gz = t % p.zgroups; t /= p.zgroups;
    int at = t % p.atiles;
    int bt = t / p.atiles;
    u.b0 = bt * p.RB; u.a0 = at * NTA;
    u.kx0 = gz * p.nshift;
    u.ns = min(p.nshift, p.K - u.kx0);
    return u;
}

template <int NTA>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(WgP p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    constexpr uint32_t LBO_B = NTA * 16;
    constexpr uint32_t G_HALF = 4 * LBO_A, G_SLOT = 2 * G_HALF;          // hi image, lo image
    constexpr uint32_t X_HALF = 4 * LBO_B, X_SLOT = 2 * X_HALF;
    constexpr uint32_t OFF_G = 0, OFF_X = OFF_G + GS * G_SLOT, OFF_BAR = OFF_X + XS * X_SLOT, OFF_SLOT = OFF_BAR + 256;
    constexpr uint32_t ACC_STRIDE = 3 * NTA;                             // TMEM columns of one accumulator set (ky-major)
    constexpr uint32_t TMEM_COLS = (2 * ACC_STRIDE <= 256) ? 256 : 512;

    const uint32_t bar0 = base + OFF_BAR;
    auto BAR_G_FULL = [&](int s) { return bar0 + 8u * s; };
    auto BAR_G_EMPTY = [&](int s) { return bar0 + 8u * (GS + s); };
    auto BAR_X_FULL = [&](int s) { return bar0 + 8u * (2 * GS + s); };
    auto BAR_X_EMPTY = [&](int s) { return bar0 + 8u * (2 * GS + XS + s); };
    auto BAR_ACC_FULL = [&](int s) { return bar0 + 8u * (2 * GS + 2 * XS + s); };
    auto BAR_ACC_EMPTY = [&](int s) { return bar0 + 8u * (2 * GS + 2 * XS + 2 + s); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < GS; ++s) { mbar_init(BAR_G_FULL(s), WG_GROUP_WARPS); mbar_init(BAR_G_EMPTY(s), 1); }
        for (int s = 0; s < XS; ++s) { mbar_init(BAR_X_FULL(s), WG_GROUP_WARPS); mbar_init(BAR_X_EMPTY(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(BAR_ACC_FULL(s), 1); mbar_init(BAR_ACC_EMPTY(s), WG_CONS_WARPS); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(base + OFF_SLOT, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + OFF_SLOT);

    // CTA -> (tile, strip partition): consecutive CTAs take DIFFERENT tiles of the SAME partition and walk its strips in the same
    // order, so the CTAs that are resident together read the same X / G rows at about the same time (L2 hits instead of
    // DRAM re-reads: every strip is needed by btiles*zgroups + atiles tiles).
    const int ntiles = p.btiles * p.atiles * p.zgroups;
    const int my_tile = blockIdx.x % ntiles, my_part = blockIdx.x / ntiles;
    const int unit_beg = my_tile * p.S + min(p.S, my_part * p.units_per_cta);
    const int unit_end = my_tile * p.S + min(p.S, (my_part + 1) * p.units_per_cta);
    const int K = p.K;

    if (warp < WG_PROD_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        // ===== MMA issuer (one thread)
        if (warp == 0 && elect_one()) {
            const uint32_t idesc = umma_idesc_tf32(128, NTA, 0, 0);
            const uint64_t a_word = ((uint64_t)(8u | (1u << 14)) << 32) | ((uint64_t)(LBO_A >> 4) << 16);   // SBO 128 B, LBO
            const uint64_t b_word = ((uint64_t)(8u | (1u << 14)) << 32) | ((uint64_t)(LBO_B >> 4) << 16);
            constexpr uint32_t a_ks = 2 * (LBO_A >> 4), b_ks = 2 * (LBO_B >> 4);
            // The issue loop is scalar code of ONE thread: it is kept to a few dozen instructions per image row (ring positions
            // are masks / shifts of running counters), otherwise it -- not the tensor core -- sets the pace.
            // A converter group owns the tasks of one parity of the global task position (= X-ring position) and its own half of the
            // G slots (slot = group + 2 * (its G-row count % (GS/2))): every EMPTY barrier has one waiting group, which meets the uses
            // of a slot in program order (see wgrad_tma.cu for what happened when ownership flipped between the groups).
            uint32_t gn0 = 0u, gn1 = 0u;            // G rows converted so far by group 0 / 1
            uint32_t xq = 0, sc = 0;                // running X-row (= task) and strip counters
            const uint32_t g0_16 = (base + OFF_G) >> 4, x0_16 = (base + OFF_X) >> 4;
            const int per_tile_strips = p.ustrips;  // strip index -> row strip: (strip / ustrips) % rstrips
            for (int unit = unit_beg; unit < unit_end; ++unit) {
                const int strip = unit - my_tile * p.S;
                const int r0 = ((strip / per_tile_strips) % p.rstrips) * p.RR;
                const int rows = min(p.RR, p.HB - r0);
                const uint32_t buf = sc & 1;
                mbar_wait_spin(BAR_ACC_EMPTY(buf), ((sc >> 1) & 1) ^ 1);
                const uint32_t d0 = tmem_base + buf * ACC_STRIDE;
                for (int j = 0; j < K - 1; ++j) mbar_wait_spin(BAR_X_FULL((xq + j) & (XS - 1)), ((xq + j) / XS) & 1);
                for (int i = 0; i < rows; ++i) {
                    const uint32_t xlast = xq + K - 1;
                    const uint32_t og = xlast & 1u, gcount = og ? gn1 : gn0;                 // the group that converted this row's task
                    const uint32_t gslot = og + 2u * (gcount % (GS / 2));
                    mbar_wait_spin(BAR_G_FULL(gslot), (gcount / (GS / 2)) & 1);
                    mbar_wait_spin(BAR_X_FULL(xlast & (XS - 1)), (xlast / XS) & 1);          // X rows i .. i+K-2 were waited for earlier
                    tc_fence_after();
                    const uint64_t g_hi = a_word + (g0_16 + gslot * (G_SLOT >> 4)), g_lo = g_hi + (G_HALF >> 4);
                    const uint32_t accf = i > 0 ? 1u : 0u;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        if (ky >= K || (p.dbg & 1)) continue;
                        const uint64_t x_hi = b_word + (x0_16 + ((xq + ky) & (XS - 1)) * (X_SLOT >> 4)), x_lo = x_hi + (X_HALF >> 4);
                        const uint32_t d = d0 + (uint32_t)ky * NTA;
                        if (p.nprod == 3) {
                            umma_tf32(d, g_hi, x_hi, idesc, accf);
                            umma_tf32(d, g_hi, x_lo, idesc, 1u);
                            umma_tf32(d, g_lo, x_hi, idesc, 1u);
                            umma_tf32(d, g_hi + a_ks, x_hi + b_ks, idesc, 1u);
                            umma_tf32(d, g_hi + a_ks, x_lo + b_ks, idesc, 1u);
                            umma_tf32(d, g_lo + a_ks, x_hi + b_ks, idesc, 1u);
                        } else {
                            umma_tf32(d, g_hi, x_hi, idesc, accf);
                            umma_tf32(d, g_hi + a_ks, x_hi + b_ks, idesc, 1u);
                        }
                    }
                    umma_commit(BAR_G_EMPTY(gslot));
                    umma_commit(BAR_X_EMPTY(xq & (XS - 1)));                           // X row i is not needed by later G rows
                    gn0 += og ^ 1u; gn1 += og; ++xq;
                }
                for (int j = 0; j < K - 1; ++j) umma_commit(BAR_X_EMPTY((xq + j) & (XS - 1)));   // the strip's bottom halo rows
                xq += K - 1;
                umma_commit(BAR_ACC_FULL(buf));
                ++sc;
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ===== converter warps: fill the rings for strip s, then drain strip s-1 from TMEM into registers
        const int ct = threadIdx.x - WG_PROD_WARPS * 32;   // 0..255
        const int q = warp & 3;                            // TMEM lane quarter
        const int half = (warp - WG_PROD_WARPS) >> 2;      // which half of the accumulator columns this warp owns (= its converter group)
        constexpr int HN = NTA / 2;                     // input channels per thread and ky
        constexpr int HC = 3 * HN;                      // accumulator registers per thread
        float acc[HC];
#pragma unroll
        for (int j = 0; j < HC; ++j) acc[j] = 0.f;
        uint32_t sc = 0;
        bool pend = false;
        int pend_unit = 0, pend_rows = 0;
        uint32_t pend_sc = 0;

        // One converter task = local X row j of a unit (image row r0 - pad_y + j) plus, for j >= K-1, G row j-(K-1).
        // The eight converter warps form TWO groups of four that take the tasks alternately.  A group's iteration is
        //   wait for its loads -> split / st.shared -> fence.proxy.async -> arrive -> issue the loads of its next task
        // i.e. the proxy fence (which also waits for the thread's outstanding global loads) always runs BEFORE new loads are
        // issued, and the global/L2 latency of one group overlaps with the conversion work of the other.
        // Everything that does not change within a unit (this thread's source pointers, scales, validity) is worked out
        // once per unit, so that a task costs a handful of address adds and six 128-bit loads per thread.
        constexpr int XI = NTA * 4 / WG_GROUP_THREADS;   // X items per thread (2 for NTA = 64, 1 for 32)
        constexpr int GI = 512 / WG_GROUP_THREADS;       // G items per thread (4)
        static_assert(XI >= 1 && NTA * 4 % WG_GROUP_THREADS == 0, "X items must tile the group");
        const int grp = (warp - WG_PROD_WARPS) >> 2;     // 0 / 1: also the half of the accumulator columns this warp owns
        const int gt = ct & (WG_GROUP_THREADS - 1);      // thread index inside the group
        struct Regs { float4 x[XI]; float4 g[GI]; float xs[XI], gs[GI]; };
        struct Cursor {                                  // position in the CTA's task sequence + this thread's load plan
            int unit, j, ntask, tile;
            uint32_t xc, gc;                             // ring positions of this task's X row / G row
            const float* xp[XI]; const float* gp[GI];
            float xs[XI], gs[GI];
            int xmode[XI], gmode[GI];                    // 0 = zero item, 1 = aligned 128-bit load, 2 = bounds-checked scalar loads
            int gx0[GI], v0, y0;
        };
        auto plan_unit = [&](Cursor& t) {
            const Unit u = decode_unit(t.unit, p, NTA);
            t.ntask = u.rows + K - 1; t.tile = u.tile; t.v0 = u.r0 - p.pad_y; t.y0 = u.r0 - (K - 1);
#pragma unroll
            for (int k = 0; k < XI; ++k) {
                const int id = gt + k * WG_GROUP_THREADS;
                const int c = item_chunk(id), row = item_row(id);
                const int ac = u.a0 + row, col = u.u0 + 4 * c;
                t.xmode[k] = (ac < p.A && col < p.WA) ? ((p.vecX && col + 3 < p.WA) ? 1 : 2) : 0;
                t.xp[k] = p.X + ((size_t)u.n * p.A + (t.xmode[k] ? ac : 0)) * p.HA * p.WA + col;
                t.xs[k] = (t.xmode[k] && p.xs) ? __ldg(p.xs + (size_t)u.n * p.A + ac) : 1.f;
            }
#pragma unroll
            for (int k = 0; k < GI; ++k) {
                const int id = gt + k * WG_GROUP_THREADS;                      // 512 items: (row m, chunk c)
                const int c = item_chunk(id), m = item_row(id);
                const int sft = m >> p.rb_shift, bc = u.b0 + (m & (p.RB - 1));
                const int x0 = u.u0 + 4 * c - (u.kx0 + sft - p.pad_x);         // G column of the chunk's first pixel
                const bool ok = sft < u.ns && bc < p.B && x0 + 3 >= 0 && x0 < p.WB;
                t.gmode[k] = ok ? ((p.vecG && (x0 & 3) == 0 && x0 >= 0 && x0 + 3 < p.WB) ? 1 : 2) : 0;
                t.gx0[k] = x0;
                t.gp[k] = p.G + ((size_t)u.n * p.B + (ok ? bc : 0)) * p.HB * p.WB;
                t.gs[k] = (ok && p.gs) ? __ldg(p.gs + (size_t)u.n * p.B + bc) : 1.f;
            }
        };
        auto load_task = [&](const Cursor& t, Regs& r) {
            const int v = (p.dbg & 2) ? -1000000 : t.v0 + t.j, y = t.y0 + t.j;
#pragma unroll
            for (int k = 0; k < XI; ++k) {
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t.xmode[k] != 0 && v >= 0 && v < p.HA) {
                    const float* src = t.xp[k] + (size_t)v * p.WA;
                    if (t.xmode[k] == 1) {
                        val = __ldg(reinterpret_cast<const float4*>(src));
                    } else {
                        const int col = (int)((t.xp[k] - p.X) % p.WA);
                        val.x = __ldg(src);
                        if (col + 1 < p.WA) val.y = __ldg(src + 1);
                        if (col + 2 < p.WA) val.z = __ldg(src + 2);
                        if (col + 3 < p.WA) val.w = __ldg(src + 3);
                    }
                }
                r.x[k] = val; r.xs[k] = t.xs[k];         // NOT multiplied here: that would wait for the load right away
            }
#pragma unroll
            for (int k = 0; k < GI; ++k) {
                float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t.gmode[k] != 0 && t.j >= K - 1 && !(p.dbg & 2)) {
                    const float* src = t.gp[k] + (size_t)y * p.WB;
                    const int x0 = t.gx0[k];
                    if (t.gmode[k] == 1) {
                        g4 = __ldg(reinterpret_cast<const float4*>(src + x0));
                    } else {
                        if (x0 >= 0) g4.x = __ldg(src + x0);
                        if (x0 + 1 >= 0 && x0 + 1 < p.WB) g4.y = __ldg(src + x0 + 1);
                        if (x0 + 2 >= 0 && x0 + 2 < p.WB) g4.z = __ldg(src + x0 + 2);
                        if (x0 + 3 < p.WB) g4.w = __ldg(src + x0 + 3);
                    }
                }
                r.g[k] = g4; r.gs[k] = t.gs[k];
            }
        };
        auto store_split = [&](uint8_t* hi_addr, uint32_t half_bytes, const float4& val, float sc_) {
            float4 h, l;
            split_tf32(val.x * sc_, h.x, l.x); split_tf32(val.y * sc_, h.y, l.y); split_tf32(val.z * sc_, h.z, l.z); split_tf32(val.w * sc_, h.w, l.w);
            *reinterpret_cast<float4*>(hi_addr) = h;
            *reinterpret_cast<float4*>(hi_addr + half_bytes) = l;
        };
        auto store_task = [&](const Cursor& t, const Regs& r) {
            const uint32_t xslot = t.xc % XS, gslot = (uint32_t)grp + 2u * (t.gc % (GS / 2));    // t.gc counts THIS GROUP's G rows
            const bool has_g = t.j >= K - 1;
            mbar_wait_spin(BAR_X_EMPTY(xslot), ((t.xc / XS) & 1) ^ 1);
            if (has_g) mbar_wait_spin(BAR_G_EMPTY(gslot), ((t.gc / (GS / 2)) & 1) ^ 1);
            uint8_t* xb = gbase + OFF_X + xslot * X_SLOT;
#pragma unroll
            for (int k = 0; k < XI; ++k) {
                if (p.dbg & 4) break;
                const int id = gt + k * WG_GROUP_THREADS;
                store_split(xb + item_chunk(id) * LBO_B + item_row(id) * 16, X_HALF, r.x[k], r.xs[k]);
            }
            if (has_g && !(p.dbg & 4)) {
                uint8_t* gb = gbase + OFF_G + gslot * G_SLOT;
#pragma unroll
                for (int k = 0; k < GI; ++k) {
                    const int id = gt + k * WG_GROUP_THREADS;
                    store_split(gb + item_chunk(id) * LBO_A + item_row(id) * 16, G_HALF, r.g[k], r.gs[k]);
                }
            }
            if (!(p.dbg & 8)) fence_proxy_async();       // one proxy fence per task, before any new global load is in flight
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(BAR_X_FULL(xslot));
                if (has_g) mbar_arrive(BAR_G_FULL(gslot));
            }
        };
        // Accumulator ownership: this thread holds, for every ky, input channels [half*HN, half*HN + HN) of TMEM lane q*32+lane.
        auto drain = [&](uint32_t k, int rows) {        // TMEM accumulator set of strip k -> registers (RN adds)
            const float kc = rz_compensation(2 * rows, p.nprod);
            const uint32_t buf = k & 1;
            mbar_wait_spin(BAR_ACC_FULL(buf), (k >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                if (ky < K) {
#pragma unroll
                    for (int cb = 0; cb < HN; cb += 16) {
                        uint32_t v[16];
                        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_STRIDE + (uint32_t)(ky * NTA + half * HN + cb), v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[ky * HN + cb + j] = fmaf(__uint_as_float(v[j]), kc, acc[ky * HN + cb + j]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR_ACC_EMPTY(buf));
        };
        auto flush = [&](const Unit& u) {               // registers -> dw (fp32 atomics; dw was zero-filled by the host)
            const int m = q * 32 + lane;
            const int s = m >> p.rb_shift, bc = u.b0 + (m & (p.RB - 1));
            const bool row_ok = s < u.ns && bc < p.B;
            int kx = u.kx0 + s;
            if (p.flip_w) kx = K - 1 - kx;
            const int ac0 = u.a0 + half * HN;
            const int nvalid = row_ok ? min(HN, p.A - ac0) : 0;
            const size_t astep = p.out_layout ? (size_t)p.B * K * K : (size_t)K * K;     // dw stride of one input channel
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int kyo = p.flip_w ? K - 1 - ky : ky;
                float* dst = p.dw + (p.out_layout ? ((((size_t)ac0 * p.B + bc) * K + kyo) * K + kx)
                                                  : ((((size_t)bc * p.A + ac0) * K + kyo) * K + kx));
#pragma unroll
                for (int j = 0; j < HN; ++j) {
                    if (ky < K && j < nvalid) atomicAdd(dst, acc[ky * HN + j]);
                    dst += astep;
                    acc[ky * HN + j] = 0.f;
                }
            }
        };
        // Both groups drain their half of the accumulator columns at the end of every unit (strip): the unit that finished
        // one strip earlier is complete in TMEM by then.  All units of a CTA belong to one tile, so there is one flush.
        auto end_of_unit = [&](int unit, int rows) {
            if (pend) drain(pend_sc, pend_rows);
            pend = true; pend_unit = unit; pend_rows = rows; pend_sc = sc;
            ++sc;
        };
        // Group g converts the local tasks j = g, g+2, ... of every unit; the ring positions follow from per-unit bases, so the
        // per-task bookkeeping is three adds.  The loads of task j+2 are issued right after task j has been published.
        Cursor C;
        uint32_t xbase = 0, gn = 0;                      // ring position of the unit's first X row (= task); G rows this group has converted
        for (int unit = unit_beg; unit < unit_end; ++unit) {
            C.unit = unit;
            plan_unit(C);
            const int ntask = C.ntask;
            Regs R;
            const int j0 = (int)((xbase ^ (uint32_t)grp) & 1u);       // the tasks whose GLOBAL position has this group's parity
            C.j = j0;
            if (C.j < ntask) load_task(C, R);
            for (int j = j0; j < ntask; j += 2) {
                C.j = j; C.xc = xbase + (uint32_t)j; C.gc = gn;
                store_task(C, R);
                if (j >= K - 1) ++gn;
                C.j = j + 2;
                if (C.j < ntask) load_task(C, R);
            }
            end_of_unit(unit, ntask - (K - 1));
            xbase += (uint32_t)ntask;
        }
        if (pend) { drain(pend_sc, pend_rows); flush(decode_unit(pend_unit, p, NTA)); }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int NTA>
int launch_wgrad(const WgP& p, int grid, cudaStream_t st) {
    constexpr uint32_t LBO_B = NTA * 16;
    const size_t smem = GS * 2 * 4 * LBO_A + XS * 2 * 4 * LBO_B + 256 + 16 + 128;
    static std::atomic<uint64_t> attr_set{0};           // one bit per device
    if (!gg::done_on_this_device(attr_set)) {
        GG_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<NTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gg::mark_done_on_this_device(attr_set);
    }
    wd_arm();
    wgrad_tc_kernel<NTA><<<grid, WG_THREADS, smem, st>>>(p);
    return gg::check_launch("conv2d_wgrad(tc)");
}

}  // namespace

namespace gg {

bool wgrad_tc_eligible(int N, int A, int HA, int WA, int B, int HB, int WB, int KH, int KW, int stride, int pad_y, int pad_x) {
    if (stride != 1 || KH != KW || KH < 1 || KH > 3) return false;
    if (pad_y > KH - 1 || pad_x > KW - 1) return false;
    if (N < 1 || A < 1 || B < 1) return false;
    if ((int64_t)A * B < 256) return false;            // 3x3-channel corner cases: nothing to gain
    if ((int64_t)N * HB * WB < 256) return false;      // 4x4 maps: launch-latency bound either way, keep the exact FFMA kernel
    (void)HA; (void)WA;
    return true;
}

// wgrad_tma.cu: the TMA-fed kernel (16-byte aligned operands)
bool wgrad_tma_eligible(const float* a, const float* b, int WA, int WB);
int wgrad_tma(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int HB, int WB, int K,
              int pad_y, int pad_x, int flip_w, int out_layout, const float* a_scale, const float* b_scale, int nprod,
              int pm_dim, unsigned pm_dead, cudaStream_t st);

int wgrad_tc(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int HB, int WB, int K, int /*KW*/,
             int pad_y, int pad_x, int flip_w, int out_layout, const float* a_scale, const float* b_scale, int nprod, int pm_dim,
             unsigned pm_dead, cudaStream_t st) {
    if (wgrad_tma_eligible(a, b, WA, WB))
        return wgrad_tma(a, b, dw, N, A, HA, WA, B, HB, WB, K, pad_y, pad_x, flip_w, out_layout, a_scale, b_scale, nprod, pm_dim, pm_dead, st);
    WgP p{};
    p.X = a; p.G = b; p.dw = dw; p.xs = a_scale; p.gs = b_scale;
    p.N = N; p.A = A; p.HA = HA; p.WA = WA; p.B = B; p.HB = HB; p.WB = WB; p.K = K; p.pad_y = pad_y; p.pad_x = pad_x;
    p.flip_w = flip_w; p.out_layout = out_layout;
    p.RB = B > 64 ? 128 : (B > 32 ? 64 : 32);
    p.rb_shift = p.RB == 128 ? 7 : (p.RB == 64 ? 6 : 5);
    p.nshift = 128 / p.RB < K ? 128 / p.RB : K;
    p.zgroups = (K + p.nshift - 1) / p.nshift;
    const int NTA = A > 32 ? 64 : 32;
    p.btiles = (B + p.RB - 1) / p.RB;
    p.atiles = (A + NTA - 1) / NTA;
    p.RR = HB < 16 ? HB : 16;   // rows per strip = TMEM chain length / 6 (96 truncating accumulates at most)
    p.ustrips = (WA + UW - 1) / UW;
    p.rstrips = (HB + p.RR - 1) / p.RR;
    const int64_t S = (int64_t)N * p.ustrips * p.rstrips;
    const int64_t total = S * p.btiles * p.atiles * p.zgroups;
    if (total > 0x7fffffffLL) { set_error("conv2d_wgrad(tc): too many work units"); return GG_EINVAL; }
    p.S = (int)S; p.total_units = (int)total;
    const int ntiles = p.btiles * p.atiles * p.zgroups;
    int64_t parts = (2LL * GG_NUM_SMS + ntiles - 1) / ntiles;          // ~2 waves of CTAs; one atomic flush per CTA
    if (parts > S) parts = S;
    if (parts < 1) parts = 1;
    p.units_per_cta = (int)((S + parts - 1) / parts);                  // strips per CTA
    parts = (S + p.units_per_cta - 1) / p.units_per_cta;
    const int grid = (int)(parts * ntiles);
    p.nprod = (nprod == GG_PREC_TF32X1) ? 1 : 3;
    p.dbg = 0;
    p.vecX = ((reinterpret_cast<uintptr_t>(a) & 15) == 0 && WA % 4 == 0) ? 1 : 0;
    p.vecG = ((reinterpret_cast<uintptr_t>(b) & 15) == 0 && WB % 4 == 0) ? 1 : 0;
    GG_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)A * B * K * K, st));
    if (NTA == 64) return launch_wgrad<64>(p, grid, st);
    return launch_wgrad<32>(p, grid, st);
}

}  // namespace gg
