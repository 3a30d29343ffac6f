// tcgen05 / TMEM weight-gradient kernel, TMA-fed, G operand in tensor memory (the path every aligned layer takes).
//
// Same contraction, GEMM view, work units and accumulation scheme as wgrad_tc.cu (read its header first):
//
//   dw[b,a,ky,kx] = sum_{n,y,x} (gs[n,b] * G[n,b,y,x]) * (xs[n,a] * X[n,a,y+ky-py,x+kx-px])
//   D_ky[m = (kx slot, grad channel b), n = input channel a] += sum_{pixels} Gshift[m, pix] * X_ky[n, pix]
//
// What is different is how the operands reach the tensor core.  ncu on wgrad_tc_kernel showed (1) the converter warps
// setting the pace (~530 warp instructions per 16-pixel row task: address arithmetic, bounds tests, global-load latency with
// two tasks in flight per SM, against ~160 that the hi/lo split needs) and, once that was fixed, (2) the shared-memory port
// saturated: 18 SS-mode MMAs per task read 108 KB of operands, the converters move another 60 KB.  Here
//
//   * one producer thread issues TMA box loads {16 (+12 apron for G) pixels, 1 row, NTA or RB channels, 1 sample} of the raw
//     fp32 rows into a staging ring.  Image borders, channel tails and the zero padding are TMA out-of-bounds zero fill:
//     no bounds test is left in the kernel.  (A box has to start on a 16-byte boundary of global memory, so the kx shift of
//     the G operand is applied by the converters when they pick their 16 pixels out of the apron row.)
//   * the G operand (M = 128 rows) lives in TENSOR MEMORY: converter thread m owns MMA row m = TMEM lane m, reads its staged row
//     (row pitch 112 B: conflict-free 128-bit reads), scales, splits into tf32 hi/lo and writes 2 x 16 columns with tcgen05.st;
//     the MMAs are issued in the TS form (A from TMEM), so only the 64 X rows of an MMA cross the shared-memory port:
//     36 KB instead of 108 KB per task, and no st.shared for G at all;
//   * the X operand (N = NTA rows) is converted into the K-major UMMA shared-memory image as before, with a rotated
//     chunk <-> lane mapping that makes both the staged read and the image stores bank-conflict free;
//   * structurally dead taps of the phase-major stride-2 weights (conv2d_resample.py: 7 of the 16 (phase, tap) blocks are
//     zero by construction) are skipped per (tile, ky): no MMAs, no drain, no flush; a tile without a live tap exits.
//
// Requirements (checked by wgrad_tma_eligible, otherwise wgrad_tc.cu's kernel runs): 16-byte aligned base pointers and row
// pitches (WA % 4 == 0, WB % 4 == 0) -- the TMA global-stride rule.
#define GG_TU_TAG 4
#include "tc_common.cuh"
#include <stdlib.h>
#include <limits.h>

using namespace ggtc;

namespace {

constexpr int CONS_WARPS = 8;
constexpr int GROUP_WARPS = 4;                          // the converter warps work as two groups that alternate tasks
constexpr int GROUP_THREADS = GROUP_WARPS * 32;
constexpr int PROD_WARPS = 4;                           // w0 = MMA issuer + TMEM owner, w1 / w2 = TMA producers (X / G rows), w3 idle (one warpgroup for setmaxnreg)
constexpr int THREADS = (PROD_WARPS + CONS_WARPS) * 32;
constexpr int UW = 16;                                  // image columns per strip (two K=8 steps)
constexpr int XS = 8;                                   // X-row ring slots in shared memory (k live rows + rows in flight)
constexpr int ST = 6;                                   // staging slots (raw fp32 rows in flight from TMA)
constexpr uint32_t ROW_B = UW * 4;                      // bytes of one staged X row (16 pixels)
constexpr int GW = UW + 12;                             // staged G row: pixels u0-4 .. u0+23; 20 are needed (every kx shift), 28 give a
constexpr uint32_t GROW_B = GW * 4;                     // row pitch of 7 x 16 B: consecutive rows fall into different bank groups
constexpr uint32_t STG_G = 128 * GROW_B;                // staged G rows (up to 128 channels), then the X rows

struct WtP {
    float* dw; const float* xs; const float* gs;
    int N, A, HA, WA, B, HB, WB, K, pad_y, pad_x, flip_w, out_layout;
    int RB, rb_shift, nshift, zgroups, btiles, atiles;
    int RR, ustrips, rstrips, S;
    int units_per_cta, nprod;
    int pm_side, pm_group;      // phase-major operand: 0 none, 1 = A channels, 2 = B channels; channels per phase group
    uint32_t pm_dead;           // bit (group*4 + ky_out*2 + kx_out): that tap of that phase group is structurally zero
};

// 8 consecutive lanes handle 8 consecutive rows of ONE 16-byte chunk (a conflict-free 128-byte quarter-warp store into the
// UMMA image), the four quarter-warps take the four chunks.  Item id + 128 = same chunk, row + 32.
__device__ __forceinline__ int item_chunk(int id) { return (id >> 3) & 3; }
__device__ __forceinline__ int item_row(int id) { return (id & 7) | ((id >> 5) << 3); }

// potentially-blocking wait (the hardware suspends the thread for a while): used by the many converter threads, which
// would otherwise hammer the shared-memory pipe that the MMA operands and the barrier traffic of the single-thread roles need
__device__ __forceinline__ void mbar_wait_block(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }

// 128-bit shared-memory read that the compiler may not split into narrower (bank-conflicting) accesses
__device__ __forceinline__ float4 lds128(const void* ptr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(ptr)));
    return v;
}

struct Tile { int b0, a0, kx0, ns; uint32_t ky_live; };
struct Strip { int n, r0, rows, u0; };

__device__ __forceinline__ Tile decode_tile(int tile, const WtP& p, int NTA) {
    Tile t;
    int gz = tile % p.zgroups; tile /= p.zgroups;
    int at = tile % p.atiles;
    int bt = tile / p.atiles;
    t.b0 = bt * p.RB; t.a0 = at * NTA;
    t.kx0 = gz * p.nshift;
    t.ns = min(p.nshift, p.K - t.kx0);
    // which ky of this tile have at least one live (phase group, kx slot) combination?
    t.ky_live = 0;
    for (int ky = 0; ky < p.K; ++ky) {
        bool live = p.pm_side == 0;
        if (!live) {
            const int c0 = p.pm_side == 1 ? t.a0 : t.b0;
            const int c1 = p.pm_side == 1 ? min(t.a0 + NTA, p.A) : min(t.b0 + p.RB, p.B);
            const int kyo = p.flip_w ? p.K - 1 - ky : ky;
            for (int g = c0 / p.pm_group; g <= (c1 - 1) / p.pm_group && g < 4; ++g)
                for (int s = 0; s < t.ns; ++s) {
                    const int kx = t.kx0 + s, kxo = p.flip_w ? p.K - 1 - kx : kx;
                    if (!((p.pm_dead >> (g * 4 + kyo * 2 + kxo)) & 1u)) live = true;
                }
        }
        if (live) t.ky_live |= 1u << ky;
    }
    return t;
}

__device__ __forceinline__ Strip decode_strip(int strip, const WtP& p) {
    Strip s;
    int us = strip % p.ustrips; strip /= p.ustrips;
    int rs = strip % p.rstrips;
    s.n = strip / p.rstrips;
    s.u0 = us * UW;
    s.r0 = rs * p.RR;
    s.rows = min(p.RR, p.HB - s.r0);
    return s;
}

template <int NTA>
__global__ void __launch_bounds__(THREADS, 1) wgrad_tma_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap gmap, WtP p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // X image in shared memory: [hi | lo][16-byte pixel chunk][ring slot][channel row] x 16 B.  Rows of CONSECUTIVE ring slots are
    // contiguous inside a chunk, so one UMMA descriptor with N = K*NTA rows covers the X rows of all K filter rows ky at once
    // (their accumulators D_ky are adjacent TMEM column ranges): 6 wide MMAs per task instead of 18 narrow ones.
    constexpr uint32_t X_ROWS = NTA * 16;                                // one ring slot inside a chunk
    // G-row slots in tensor memory (32 columns each: 16 hi + 16 lo).  The slot ring is a latency loop (converter -> tcgen05.st ->
    // barrier -> MMA issue -> tensor queue -> commit -> barrier -> converter, ~3 600 cycles): with 4 slots every shape ran at
    // ~900 cycles per row task whatever its MMA time.  NTA = 32 has the TMEM columns for 8 slots; at NTA = 64 the two
    // accumulator sets (2 x 192 columns) leave room for 4.
    constexpr int GS = NTA == 32 ? 8 : 4;
    constexpr uint32_t LBO_B = XS * X_ROWS;                              // chunk pitch
    constexpr uint32_t X_HALF = 4 * LBO_B;                               // hi image, then lo image
    constexpr uint32_t STG_SLOT = STG_G + NTA * ROW_B;                   // 18 KB (NTA = 64) / 16 KB (NTA = 32)
    constexpr uint32_t OFF_X = 0, OFF_STG = OFF_X + 2 * X_HALF, OFF_BAR = OFF_STG + ST * STG_SLOT, OFF_SLOT = OFF_BAR + 512;
    static_assert(OFF_STG % 128 == 0 && STG_SLOT % 128 == 0 && STG_G % 128 == 0, "TMA destinations must be 128-byte aligned");
    constexpr uint32_t ACC_STRIDE = 3 * NTA;                             // TMEM columns of one accumulator set (ky-major)
    constexpr uint32_t TM_G = 2 * ACC_STRIDE;                            // first column of the G slots
    constexpr uint32_t TMEM_COLS = 512;                                  // 2 accumulator sets + GS x 32 columns of G (384 + 128 at NTA = 64)
    static_assert(TM_G + GS * 32 <= TMEM_COLS, "tensor memory budget");

    const uint32_t bar0 = base + OFF_BAR;
    auto BAR_G_FULL = [&](int s) { return bar0 + 8u * s; };
    auto BAR_G_EMPTY = [&](int s) { return bar0 + 8u * (GS + s); };
    auto BAR_X_FULL = [&](int s) { return bar0 + 8u * (2 * GS + s); };
    auto BAR_X_EMPTY = [&](int s) { return bar0 + 8u * (2 * GS + XS + s); };
    auto BAR_ACC_FULL = [&](int s) { return bar0 + 8u * (2 * GS + 2 * XS + s); };
    auto BAR_ACC_EMPTY = [&](int s) { return bar0 + 8u * (2 * GS + 2 * XS + 2 + s); };
    auto BAR_STG_FULL = [&](int s) { return bar0 + 8u * (2 * GS + 2 * XS + 4 + s); };
    auto BAR_STG_EMPTY = [&](int s) { return bar0 + 8u * (2 * GS + 2 * XS + 4 + ST + s); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < GS; ++s) { mbar_init(BAR_G_FULL(s), GROUP_WARPS); mbar_init(BAR_G_EMPTY(s), 1); }
        for (int s = 0; s < XS; ++s) { mbar_init(BAR_X_FULL(s), GROUP_WARPS); mbar_init(BAR_X_EMPTY(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(BAR_ACC_FULL(s), 1); mbar_init(BAR_ACC_EMPTY(s), CONS_WARPS); }
        for (int s = 0; s < ST; ++s) { mbar_init(BAR_STG_FULL(s), 2); mbar_init(BAR_STG_EMPTY(s), GROUP_WARPS); }   // FULL: one arrive per producer
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(base + OFF_SLOT, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + OFF_SLOT);

    // CTA -> (tile, strip partition): consecutive CTAs take DIFFERENT tiles of the SAME partition and walk its strips in the same
    // order, so the CTAs that are resident together read the same X / G rows at about the same time (L2 hits).
    const int ntiles = p.btiles * p.atiles * p.zgroups;
    const int my_tile = blockIdx.x % ntiles, my_part = blockIdx.x / ntiles;
    const Tile T = decode_tile(my_tile, p, NTA);
    const int strip_beg = T.ky_live ? min(p.S, my_part * p.units_per_cta) : 0;
    const int strip_end = T.ky_live ? min(p.S, (my_part + 1) * p.units_per_cta) : 0;     // a tile with no live tap has nothing to do
    const int K = p.K;

    if (warp < PROD_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0 && elect_one()) {
            // ===== MMA issuer (one thread): D_ky[tmem] += G[tmem] * X_ky[smem]
            const uint32_t idesc = umma_idesc_tf32(128, NTA, 0, 0);
            const uint32_t idesc_w = umma_idesc_tf32(128, K * NTA, 0, 0);          // all ky in one instruction
            const bool all_live = T.ky_live == (1u << K) - 1u;
            const uint64_t b_word = ((uint64_t)(8u | (1u << 14)) << 32) | ((uint64_t)(LBO_B >> 4) << 16);   // SBO 128 B, LBO
            constexpr uint32_t b_ks = 2 * (LBO_B >> 4);
            // Ring positions.  A converter GROUP owns the tasks of one parity of the global task position (= X-ring position), and
            // with them: the X slots of that parity, and its own half of the G slots (slot = group + 2 * (its G-row count % (GS/2))).
            // Every EMPTY barrier therefore has ONE waiting group, which meets the uses of a slot in program order.  (With the tasks
            // dealt out by their index INSIDE a strip, slot ownership flipped between the groups whenever a strip had an odd number
            // of tasks (K = 2) or of rows (HB = 16 n + 1): a group could then wait for release #u-1 of a slot before release #u-2 had
            // happened, which an mbarrier parity wait cannot tell apart from success -- overwritten operands, and with the barrier
            // over-arrived, launch failures.  Rare at 3 MMAs per product, immediate at 1.)
            uint32_t gn0 = 0u, gn1 = 0u;            // G rows converted so far by group 0 / 1 (scalars: an indexed array would live on the stack)
            uint32_t xq = 0, sc = 0;                // running X-row (= task) and strip counters
            const uint32_t x0_16 = (base + OFF_X) >> 4;
            const bool three = p.nprod == 3;
            for (int strip = strip_beg; strip < strip_end; ++strip) {
                const int r0 = ((strip / p.ustrips) % p.rstrips) * p.RR;
                const int rows = min(p.RR, p.HB - r0);
                const uint32_t buf = sc & 1;
                mbar_wait_spin(BAR_ACC_EMPTY(buf), ((sc >> 1) & 1) ^ 1);
                const uint32_t d0 = tmem_base + buf * ACC_STRIDE;
                for (int j = 0; j < K - 1; ++j) mbar_wait_spin(BAR_X_FULL((xq + j) & (XS - 1)), ((xq + j) / XS) & 1);
                for (int i = 0; i < rows; ++i) {
                    const uint32_t xlast = xq + K - 1, xfirst = xq & (XS - 1);
                    const uint32_t og = xlast & 1u;                                          // the group that converted this row's task
                    const uint32_t gcount = og ? gn1 : gn0;
                    const uint32_t gslot = og + 2u * (gcount % (GS / 2));
                    mbar_wait_spin(BAR_G_FULL(gslot), (gcount / (GS / 2)) & 1);
                    mbar_wait_spin(BAR_X_FULL(xlast & (XS - 1)), (xlast / XS) & 1);          // X rows i .. i+K-2 were waited for earlier
                    tc_fence_after();
                    const uint32_t g_hi = tmem_base + TM_G + gslot * 32, g_lo = g_hi + 16;   // columns: 16 pixels hi, 16 pixels lo
                    const uint32_t accf = i > 0 ? 1u : 0u;
                    if (all_live && xfirst + K <= XS) {
                        // the K ring slots of this row's window are consecutive in memory: one N = K*NTA instruction per product
                        const uint64_t x_hi = b_word + (x0_16 + xfirst * (X_ROWS >> 4)), x_lo = x_hi + (X_HALF >> 4);
                        umma_tf32_ts(d0, g_hi, x_hi, idesc_w, accf);
                        if (three) { umma_tf32_ts(d0, g_hi, x_lo, idesc_w, 1u); umma_tf32_ts(d0, g_lo, x_hi, idesc_w, 1u); }
                        umma_tf32_ts(d0, g_hi + 8, x_hi + b_ks, idesc_w, 1u);
                        if (three) { umma_tf32_ts(d0, g_hi + 8, x_lo + b_ks, idesc_w, 1u); umma_tf32_ts(d0, g_lo + 8, x_hi + b_ks, idesc_w, 1u); }
                    } else {
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            if (ky >= K || !((T.ky_live >> ky) & 1u)) continue;
                            const uint64_t x_hi = b_word + (x0_16 + ((xq + ky) & (XS - 1)) * (X_ROWS >> 4)), x_lo = x_hi + (X_HALF >> 4);
                            const uint32_t d = d0 + (uint32_t)ky * NTA;
                            umma_tf32_ts(d, g_hi, x_hi, idesc, accf);
                            if (three) { umma_tf32_ts(d, g_hi, x_lo, idesc, 1u); umma_tf32_ts(d, g_lo, x_hi, idesc, 1u); }
                            umma_tf32_ts(d, g_hi + 8, x_hi + b_ks, idesc, 1u);
                            if (three) { umma_tf32_ts(d, g_hi + 8, x_lo + b_ks, idesc, 1u); umma_tf32_ts(d, g_lo + 8, x_hi + b_ks, idesc, 1u); }
                        }
                    }
                    umma_commit(BAR_G_EMPTY(gslot));
                    umma_commit(BAR_X_EMPTY(xfirst));                                  // X row i is not needed by later G rows
                    gn0 += og ^ 1u; gn1 += og; ++xq;
                }
                for (int j = 0; j < K - 1; ++j) umma_commit(BAR_X_EMPTY((xq + j) & (XS - 1)));   // the strip's bottom halo rows
                xq += K - 1;
                umma_commit(BAR_ACC_FULL(buf));
                ++sc;
            }
        } else if ((warp == 1 || warp == 2) && elect_one()) {
            // ===== TMA producers (one thread per operand: warp 1 = X rows, warp 2 = G rows): raw fp32 rows of task j of every strip
            // -> staging ring (two threads so that the scalar issue code of one producer cannot set the pace of the kernel)
            const bool is_g = warp == 2;
            uint32_t tc = 0;
            const uint32_t stg0 = base + OFF_STG + (is_g ? 0u : STG_G);
            const CUtensorMap* map = is_g ? &gmap : &xmap;
            const uint32_t bytes = is_g ? (uint32_t)p.RB * GROW_B : NTA * ROW_B;
            const int ch0 = is_g ? T.b0 : T.a0;
            for (int strip = strip_beg; strip < strip_end; ++strip) {
                const Strip s = decode_strip(strip, p);
                const int ntask = s.rows + K - 1;
                // A TMA box must start on a 16-byte boundary of global memory, so the kx shift cannot be put into the box
                // coordinate: ONE aligned G box with a 4-pixel apron on the left serves all shifts (the converters pick).
                const int cx = is_g ? s.u0 - 4 : s.u0;
                const int cy = is_g ? s.r0 - (K - 1) : s.r0 - p.pad_y;
                for (int j = 0; j < ntask; ++j, ++tc) {
                    const uint32_t slot = tc % ST;
                    mbar_wait_spin(BAR_STG_EMPTY(slot), ((tc / ST) & 1) ^ 1);
                    if (is_g && j < K - 1) { mbar_arrive(BAR_STG_FULL(slot)); continue; }      // the strip's top halo rows have no G row
                    mbar_expect_tx(BAR_STG_FULL(slot), bytes);
                    tma_load_4d(stg0 + slot * STG_SLOT, map, BAR_STG_FULL(slot), cx, cy + j, ch0, s.n);
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ===== converter warps: staged fp32 rows -> tf32 hi/lo (G: tensor memory, X: UMMA shared-memory image); then drain strip s-1
        const int ct = threadIdx.x - PROD_WARPS * 32;      // 0..255
        const int q = warp & 3;                            // TMEM lane quarter this warp may access
        const int grp = (warp - PROD_WARPS) >> 2;          // converter group = half of the accumulator columns this warp owns
        const int half = grp;
        constexpr int HN = NTA / 2;                        // input channels per thread and ky
        constexpr int HC = 3 * HN;                         // accumulator registers per thread
        float acc[HC];
#pragma unroll
        for (int j = 0; j < HC; ++j) acc[j] = 0.f;

        constexpr int XI = NTA * 4 / GROUP_THREADS;        // X items per thread (2 for NTA = 64, 1 for 32)
        const int gt = ct & (GROUP_THREADS - 1);           // thread index inside the group
        // X item k of this thread: row row0 + 32 k, 16-byte chunk cex.  Within a quarter-warp (8 lanes = 8 consecutive rows) the
        // chunk index rotates with the row pair, cex = c0 ^ ((row >> 1) & 3), so that BOTH the 128-bit read of the staged
        // row-major tile (row pitch 64 B) and the 128-bit stores into the chunk-major UMMA image (row pitch 16 B) touch eight
        // different 16-byte bank groups: conflict free without a TMA swizzle.
        const int c0 = item_chunk(gt), row0 = item_row(gt);
        const int cex = c0 ^ ((row0 >> 1) & 3);
        const uint32_t xsrc0 = STG_G + (uint32_t)row0 * ROW_B + (uint32_t)(cex << 4);
        const uint32_t xdst0 = (uint32_t)cex * LBO_B + (uint32_t)row0 * 16;      // + ring slot * X_ROWS
        // G: this thread owns MMA row m = TMEM lane m = (kx slot sft, channel b).  Its 16 pixels u0 - sh .. u0 - sh + 15
        // (sh = kx - pad_x) are the staged pixels t .. t + 15 with t = 4 - sh in 2 .. 6: five aligned 128-bit reads starting at
        // chunk t >> 2, then a (warp-uniform) register pick by t & 3.
        const int m = q * 32 + lane;
        const int g_sft = m >> p.rb_shift, g_b = m & (p.RB - 1);
        const int g_t = 4 - (T.kx0 + g_sft - p.pad_x);
        const uint32_t gsrc = (uint32_t)g_b * GROW_B + (uint32_t)((g_t >> 2) << 4);
        const int g_e0 = g_t & 3;
        const bool g_zero = g_sft >= T.ns;                 // rows of kx slots this tile does not have: written as zeros
        const uint32_t g_taddr = tmem_base + ((uint32_t)(q * 32) << 16) + TM_G;
        float xsc[XI], gsc = 1.f;
#pragma unroll
        for (int k = 0; k < XI; ++k) xsc[k] = 1.f;
        int cur_n = -1;

        auto store_split = [&](uint8_t* hi_addr, uint32_t half_bytes, const float4& val, float sc_) {
            float4 h, l;
            split_tf32(val.x * sc_, h.x, l.x); split_tf32(val.y * sc_, h.y, l.y); split_tf32(val.z * sc_, h.z, l.z); split_tf32(val.w * sc_, h.w, l.w);
            *reinterpret_cast<float4*>(hi_addr) = h;
            *reinterpret_cast<float4*>(hi_addr + half_bytes) = l;
        };
        // Accumulator ownership: this thread holds, for every ky, input channels [half*HN, half*HN + HN) of TMEM lane q*32+lane.
        auto drain = [&](uint32_t k, int rows) {           // TMEM accumulator set of strip k -> registers (RN adds)
            const float kc = rz_compensation(2 * rows, p.nprod);
            const uint32_t buf = k & 1;
            mbar_wait_block(BAR_ACC_FULL(buf), (k >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                if (ky < K && ((T.ky_live >> ky) & 1u)) {
                    // all loads of one filter row are issued before one wait, so that their latencies overlap
                    uint32_t v[HN / 16][16];
#pragma unroll
                    for (int cb = 0; cb < HN; cb += 16)
                        tmem_ld16_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_STRIDE + (uint32_t)(ky * NTA + half * HN + cb), v[cb / 16]);
                    tmem_wait_ld();
#pragma unroll
                    for (int cb = 0; cb < HN; cb += 16)
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[ky * HN + cb + j] = fmaf(__uint_as_float(v[cb / 16][j]), kc, acc[ky * HN + cb + j]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR_ACC_EMPTY(buf));
        };
        auto flush = [&]() {                               // registers -> dw (fp32 atomics; dw was zero-filled by the host)
            const int s = m >> p.rb_shift, bc = T.b0 + (m & (p.RB - 1));
            const bool row_ok = s < T.ns && bc < p.B;
            int kx = T.kx0 + s;
            if (p.flip_w) kx = K - 1 - kx;
            const int ac0 = T.a0 + half * HN;
            const int nvalid = row_ok ? min(HN, p.A - ac0) : 0;
            const size_t astep = p.out_layout ? (size_t)p.B * K * K : (size_t)K * K;     // dw stride of one input channel
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int kyo = p.flip_w ? K - 1 - ky : ky;
                float* dst = p.dw + (p.out_layout ? ((((size_t)ac0 * p.B + bc) * K + kyo) * K + kx)
                                                  : ((((size_t)bc * p.A + ac0) * K + kyo) * K + kx));
                const bool ky_ok = ky < K && ((T.ky_live >> ky) & 1u);
#pragma unroll
                for (int j = 0; j < HN; ++j) {
                    if (ky_ok && j < nvalid) atomicAdd(dst, acc[ky * HN + j]);
                    dst += astep;
                }
            }
        };

        uint32_t sc = 0, tbase = 0, xbase = 0;             // strip counter; staging / X-ring positions of the strip's first task
        uint32_t gn = 0;                                   // G rows this GROUP has converted (its own half of the G-slot ring)
        bool pend = false;
        int pend_rows = 0;
        uint32_t pend_sc = 0;
        for (int strip = strip_beg; strip < strip_end; ++strip) {
            const Strip s = decode_strip(strip, p);
            const int ntask = s.rows + K - 1;
            if (s.n != cur_n) {                            // per-sample channel scales of this thread's rows (samples change slowly)
                cur_n = s.n;
#pragma unroll
                for (int k = 0; k < XI; ++k) {
                    const int ac = T.a0 + row0 + 32 * k;
                    xsc[k] = (p.xs && ac < p.A) ? __ldg(p.xs + (size_t)s.n * p.A + ac) : 1.f;
                }
                const int bc = T.b0 + g_b;
                gsc = (p.gs && bc < p.B) ? __ldg(p.gs + (size_t)s.n * p.B + bc) : 1.f;
            }
            for (int j = (int)((tbase ^ (uint32_t)grp) & 1u); j < ntask; j += 2) {     // the tasks whose GLOBAL position has this group's parity
                const uint32_t tcn = tbase + (uint32_t)j, slot = tcn % ST;
                const uint32_t xc = xbase + (uint32_t)j, xslot = xc & (XS - 1);
                const bool has_g = j >= K - 1;
                const uint32_t gslot = (uint32_t)grp + 2u * (gn % (GS / 2)), gphase = (gn / (GS / 2)) & 1u;
                mbar_wait_block(BAR_STG_FULL(slot), (tcn / ST) & 1);
                const uint8_t* stg = gbase + OFF_STG + slot * STG_SLOT;
                float4 xv[XI];
#pragma unroll
                for (int k = 0; k < XI; ++k) xv[k] = lds128(stg + xsrc0 + k * 32 * ROW_B);
                float g[16];
                if (has_g) {
                    float v[20];
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        const float4 t4 = lds128(stg + gsrc + 16 * c);
                        v[4 * c] = t4.x; v[4 * c + 1] = t4.y; v[4 * c + 2] = t4.z; v[4 * c + 3] = t4.w;
                    }
                    // warp-uniform pick: all 32 rows of a warp belong to the same kx slot
                    if (g_e0 == 0) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) g[e] = v[e];
                    } else if (g_e0 == 1) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) g[e] = v[e + 1];
                    } else if (g_e0 == 2) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) g[e] = v[e + 2];
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) g[e] = v[e + 3];
                    }
                }
                mbar_wait_block(BAR_X_EMPTY(xslot), ((xc / XS) & 1) ^ 1);
                uint8_t* xb = gbase + OFF_X + xslot * X_ROWS + xdst0;
#pragma unroll
                for (int k = 0; k < XI; ++k) store_split(xb + k * 32 * 16, X_HALF, xv[k], xsc[k]);
                if (has_g) {
                    uint32_t hi[16], lo[16];
                    const float gs_ = g_zero ? 0.f : gsc;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        float h, l;
                        split_tf32(g[e] * gs_, h, l);
                        hi[e] = __float_as_uint(h); lo[e] = __float_as_uint(l);
                    }
                    mbar_wait_block(BAR_G_EMPTY(gslot), gphase ^ 1u);
                    tc_fence_after();
                    tmem_st16(g_taddr + gslot * 32, hi);
                    tmem_st16(g_taddr + gslot * 32 + 16, lo);
                    tmem_wait_st();
                    tc_fence_before();
                }
                fence_proxy_async();                       // generic-proxy stores (X image) -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(BAR_STG_EMPTY(slot));
                    mbar_arrive(BAR_X_FULL(xslot));
                    if (has_g) mbar_arrive(BAR_G_FULL(gslot));
                }
                if (has_g) ++gn;
            }
            // both groups drain their half of the accumulator columns of the strip that finished one strip earlier
            if (pend) drain(pend_sc, pend_rows);
            pend = true; pend_rows = s.rows; pend_sc = sc;
            ++sc;
            tbase += (uint32_t)ntask; xbase += (uint32_t)ntask;
        }
        if (pend) { drain(pend_sc, pend_rows); flush(); }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int NTA>
int launch_wgrad_tma(const CUtensorMap& xmap, const CUtensorMap& gmap, const WtP& p, int grid, cudaStream_t st) {
    const size_t smem = XS * 2 * 4 * (NTA * 16) + ST * (STG_G + NTA * ROW_B) + 512 + 16 + 128;
    static std::atomic<uint64_t> attr_set{0};           // one bit per device
    if (!gg::done_on_this_device(attr_set)) {
        GG_CUDA(cudaFuncSetAttribute(wgrad_tma_kernel<NTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gg::mark_done_on_this_device(attr_set);
    }
    wd_arm();
    wgrad_tma_kernel<NTA><<<grid, THREADS, smem, st>>>(xmap, gmap, p);
    return gg::check_launch("conv2d_wgrad(tc)");
}

int encode_rows(CUtensorMap* map, const float* t, int N, int C, int H, int W, int box_w, int box_c) {
    gg::EncodeTiledFn encode = gg::get_encode_fn();
    if (!encode) { gg::set_error("conv2d_wgrad(tc): cuTensorMapEncodeTiled is unavailable"); return GG_ECUDA; }
    cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)box_w, 1, (cuuint32_t)box_c, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(t), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { gg::set_error("conv2d_wgrad(tc): cuTensorMapEncodeTiled failed (%d)", (int)cr); return GG_ECUDA; }
    return GG_OK;
}

}  // namespace

namespace gg {

bool wgrad_tma_eligible(const float* a, const float* b, int WA, int WB) {
    return (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0 && WA % 4 == 0 && WB % 4 == 0;
}

// pm_dim: 0 = no phase-major structure; 1 / 2 = dimension 0 / 1 of dw holds 4 phase groups (py, px) of equal size;
// pm_dead: bit (group*4 + ky*2 + kx) set = that tap of that group is structurally zero in dw (K == 2)
int wgrad_tma(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int HB, int WB, int K,
              int pad_y, int pad_x, int flip_w, int out_layout, const float* a_scale, const float* b_scale, int nprod,
              int pm_dim, unsigned pm_dead, cudaStream_t st) {
    WtP p{};
    p.dw = dw; p.xs = a_scale; p.gs = b_scale;
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
    if (S * p.btiles * p.atiles * p.zgroups > 0x7fffffffLL) { set_error("conv2d_wgrad(tc): too many work units"); return GG_EINVAL; }
    p.S = (int)S;
    // phase-major hint: dw dim 0 is B (out_layout 0) or A (out_layout 1)
    p.pm_side = 0; p.pm_dead = 0; p.pm_group = 1;
    if (pm_dim != 0 && K == 2 && pm_dead != 0) {
        const bool on_a = (pm_dim == 1) == (out_layout != 0);
        const int ch = on_a ? A : B;
        if (ch % 4 == 0) { p.pm_side = on_a ? 1 : 2; p.pm_group = ch / 4; p.pm_dead = pm_dead; }
    }
    const int ntiles = p.btiles * p.atiles * p.zgroups;
    // Strip partitions per tile: one CTA per SM is resident, so the kernel takes ceil(grid / 148) rounds of `strips per CTA`
    // each; pick the partition count (>= ~2 waves for balance, one atomic flush per CTA) that minimises rounds x strips.
    int64_t parts = 1, best = INT64_MAX;
    const int64_t pmin = (2LL * GG_NUM_SMS + ntiles - 1) / ntiles;
    for (int64_t c = pmin < S ? pmin : S; c <= S && c <= pmin + 40; ++c) {
        if (c < 1) continue;
        const int64_t per = (S + c - 1) / c, ctas = ((S + per - 1) / per) * ntiles;
        const int64_t cost = ((ctas + GG_NUM_SMS - 1) / GG_NUM_SMS) * (per + 1);      // +1: fixed cost of a CTA (setup, drain, flush)
        if (cost < best) { best = cost; parts = c; }
    }
    p.units_per_cta = (int)((S + parts - 1) / parts);                  // strips per CTA
    parts = (S + p.units_per_cta - 1) / p.units_per_cta;
    const int grid = (int)(parts * ntiles);
    p.nprod = (nprod == GG_PREC_TF32X1) ? 1 : 3;
    CUtensorMap xmap, gmap;
    int rc = encode_rows(&xmap, a, N, A, HA, WA, UW, NTA);
    if (rc != GG_OK) return rc;
    rc = encode_rows(&gmap, b, N, B, HB, WB, GW, p.RB);
    if (rc != GG_OK) return rc;
    GG_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)A * B * K * K, st));
    if (NTA == 64) return launch_wgrad_tma<64>(xmap, gmap, p, grid, st);
    return launch_wgrad_tma<32>(xmap, gmap, p, grid, st);
}

}  // namespace gg
