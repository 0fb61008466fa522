// sw_long.cu -- one very long Smith-Waterman alignment spread over every SM of one or several B200s.
//
// Same recurrence as sw_kernels.cu (antidiagonalSmithWaterman.c:290-335), raw-byte comparison, s32
// DPX.  The DP matrix is cut into column stripes of 32*K columns.  A warp owns one stripe at a time
// and sweeps all rows of it systolically (lane t holds K columns in registers and is one row behind
// lane t-1, exactly like sw_wave_kernel); the stripes themselves form a second, coarser wavefront:
// stripe j may process row block i as soon as stripe j-1 has published the right boundary column
// (H+goe and E per row) of that block.
//
//   - ONE boundary array per GPU (one 16-byte entry per row) is shared by all stripes and updated in
//     place: entry r always holds the boundary of the last stripe that passed row r, and stripes pass a
//     row strictly in order.
//   - An entry is {H+goe, tag, E, tag} with tag = the (global) index of the stripe that wrote it, stored
//     with ONE 128-bit store.  The consumer polls the data itself: no flags, no fences, one memory round
//     trip per hand-off.  Lane 31 stages the boundary in shared memory and the warp hands 32 rows on with
//     one coalesced store per block; polling is warp-uniform.  All warps are co-resident (cooperative
//     launch), so a poll always ends.
//   - Multi-GPU: GPU g owns a contiguous range of columns.  The last stripe of GPU g writes its entries
//     straight into GPU g+1's boundary array with peer stores over NVLink (cudaDeviceEnablePeerAccess);
//     no collective is involved, the final score is the max of the per-GPU maxima.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "common.cuh"

namespace agx {

namespace {

constexpr int LONG_WARPS = 4;     // warps per CTA
constexpr int LONG_RING = 64;

struct LongArgs {
    const uint8_t *a;         // columns owned by this GPU
    int32_t la;
    const uint8_t *b;         // rows
    int32_t lb;
    int4 *bnd;                // [lb] local boundary entries (in place)
    int4 *next_bnd;           // boundary array of the next GPU (peer) or nullptr
    int32_t *best;            // running maximum (atomicMax)
    int32_t stripe_base;      // global index of this GPU's stripe 0 (0: the true left edge of the matrix)
    SwScoring sc;
    const uint8_t *lut;       // CODED kernels: byte -> symbol code 0..6 (both sequences use <= 7 distinct bytes)
    int32_t one;              // the constant 1, opaque to ptxas: x + y as IMAD (FMA pipe) instead of IADD3 (ALU pipe)
};

// prmt.b32 in its default mode: selector nibble bit 3 replicates the sign of the selected byte
__device__ __forceinline__ int32_t prmt_s(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return (int32_t)d;
}
// a + b on the FMA pipe (IMAD) -- the ALU pipe is the bottleneck of these kernels
__device__ __forceinline__ int32_t add_fma(int32_t a, int32_t b, int32_t one)
{
    int32_t d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
}

// row steps unrolled per loop trip (measured at 1 Mbp on one GPU: 1 -> 415 ms, 2 -> 407 ms, 4 -> 420 ms)
#ifndef AGX_LONG_UNROLL
#define AGX_LONG_UNROLL 2
#endif
constexpr int LONG_UNROLL = AGX_LONG_UNROLL;

// A boundary entry may be (re)written by another SM or by a peer GPU while it is polled: relaxed (strong)
// accesses at the narrowest scope that covers writer and reader -- .gpu inside one GPU, .sys across NVLink.
// (ld/st.volatile and plain .cg stores were measured too: no difference.)  Requesting the next block's entries
// a block ahead was also tried: it needs a two-block start-up slack and came out 5-10 % slower.
__device__ __forceinline__ int4 ld_entry(const int4 *p, bool sys)
{
    int4 v;
    if (sys)
        asm volatile("ld.relaxed.sys.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    else
        asm volatile("ld.relaxed.gpu.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_entry(int4 *p, int4 v, bool sys)
{
    if (sys)
        asm volatile("st.relaxed.sys.global.v4.s32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else
        asm volatile("st.relaxed.gpu.global.v4.s32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// CODED: both sequences use at most 7 distinct bytes.  Columns carry a PRMT selector instead of their byte,
// rows an 8-byte table (substitution score - goe for each symbol code), and the substitution score of a cell
// is ONE PRMT (sign-extending byte select) instead of ISETP + SEL; the two additions of a cell go to the FMA
// pipe as IMADs.  ALU-pipe instructions per cell: 7.5 -> 4.5 (lean chain), 8.5 -> 5.5 (short chain).
template <int K, bool SHORT, bool CODED>
__global__ void __launch_bounds__(LONG_WARPS * 32)
sw_long_kernel(LongArgs g)
{
    constexpr int W = 32 * K;
    __shared__ int32_t r_byte[LONG_WARPS][LONG_RING];      // row byte, or the low half of the row's score table
    __shared__ int32_t r_hi[CODED ? LONG_WARPS : 1][CODED ? LONG_RING : 1];
    __shared__ int32_t r_g[LONG_WARPS][LONG_RING];
    __shared__ int32_t r_e[LONG_WARPS][LONG_RING];
    __shared__ int2 stage[LONG_WARPS][32];      // boundary of the rows finished in the current block

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * LONG_WARPS + wib;
    const int n_warps = gridDim.x * LONG_WARPS;
    const int32_t goe = g.sc.gap_open + g.sc.gap_extend;
    const int32_t ext = g.sc.gap_extend;
    const int32_t sub_match = g.sc.match - goe, sub_mis = g.sc.mismatch - goe;
    const int32_t lb = g.lb;
    const int n_stripes = (g.la + W - 1) / W;
    int32_t bestg = goe;          // running max of H + goe
    const int32_t one = g.one;
    const uint32_t xb4 = (uint32_t)(uint8_t)(int8_t)sub_mis * 0x01010101u;          // every symbol: mismatch
    const uint32_t mxor = (uint32_t)(uint8_t)(int8_t)sub_mis ^ (uint32_t)(uint8_t)(int8_t)sub_match;
    // row symbol -> its 8-byte score table (byte k = score against symbol code k); code 7 never matches
    auto row_table = [&](int32_t r, uint32_t &lo, uint32_t &hi) {
        lo = xb4; hi = xb4;
        if (r < lb) {
            const uint32_t c = g.lut[g.b[r]];
            if (c < 4) lo ^= mxor << (8 * c); else hi ^= mxor << (8 * (c - 4));
        }
    };

    for (int st = warp; st < n_stripes; st += n_warps) {
        const int c0 = st * W + lane * K;
        int32_t acol[K], Gp[K], F[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if constexpr (CODED) {
                const uint32_t c = (c0 + j < g.la) ? (uint32_t)g.lut[g.a[c0 + j]] : 7u;
                acol[j] = (int32_t)(c | ((8u | c) * 0x1110u));              // byte c, sign-extended to 32 bits
            } else {
                acol[j] = (c0 + j < g.la) ? (int32_t)g.a[c0 + j] : 0x100;
            }
            Gp[j] = goe;
            F[j] = goe;
        }
        int32_t g_out = goe, e_out = goe, g_in_prev = goe;
        const int32_t gst = g.stripe_base + st;                          // global stripe index = my tag
        const bool left_edge = (gst == 0);
        const bool last = (st == n_stripes - 1);
        int4 *out_bnd = last ? g.next_bnd : g.bnd;                       // nullptr: nothing to hand on
        const bool out_remote = last;                                    // the next GPU's array, over NVLink
        const bool in_remote = (st == 0);                                // written by the previous GPU
        const int S = lb + 31;

        // inputs of the first 32 rows
        int32_t nb = 0x200;
        uint32_t nhi = 0;
        int4 nx = make_int4(goe, gst - 1, goe, gst - 1);
        if constexpr (CODED) { uint32_t lo; row_table(lane, lo, nhi); nb = (int32_t)lo; }
        if (lane < lb) {
            if constexpr (!CODED) nb = g.b[lane];
            if (!left_edge) nx = ld_entry(g.bnd + lane, in_remote);
        }
        for (int s0 = 0; s0 < S; s0 += 32) {
            {
                const int r = s0 + lane;
                if (!left_edge) {
                    // wait until the left neighbour has handed these 32 rows on
                    unsigned ns = 32;
                    while (__any_sync(0xffffffffu, nx.y != gst - 1 || nx.w != gst - 1)) {
                        __nanosleep(ns);
                        if (ns < 512) ns *= 2;
                        if (r < lb && (nx.y != gst - 1 || nx.w != gst - 1)) nx = ld_entry(g.bnd + r, in_remote);
                    }
                }
                r_byte[wib][r & (LONG_RING - 1)] = nb;
                if constexpr (CODED) r_hi[wib][r & (LONG_RING - 1)] = (int32_t)nhi;
                r_g[wib][r & (LONG_RING - 1)] = nx.x;
                r_e[wib][r & (LONG_RING - 1)] = nx.z;
            }
            __syncwarp();
            const int send = min(32, S - s0);
#pragma unroll LONG_UNROLL
            for (int u = 0; u < send; ++u) {
                const int s = s0 + u;
                const int slot = (s - lane) & (LONG_RING - 1);
                int32_t rb = (s - lane >= 0) ? r_byte[wib][slot] : (CODED ? (int32_t)xb4 : 0x200);
                uint32_t rhi = xb4;
                if constexpr (CODED) { if (s - lane >= 0) rhi = (uint32_t)r_hi[wib][slot]; }
                int32_t g_in = __shfl_up_sync(0xffffffffu, g_out, 1);
                int32_t e = __shfl_up_sync(0xffffffffu, e_out, 1);
                if (lane == 0) { g_in = r_g[wib][slot]; e = r_e[wib][slot]; }
                int32_t gdiag = g_in_prev;
                g_in_prev = g_in;
                int32_t gleft = g_in;
                if constexpr (SHORT) {
                    // Short dependency chain: with T = max(F, diag + s, 0) (known from the previous row),
                    //   E[j] = max(E[j-1] + ext, T[j-1] + goe)   (E[j-1] + ext >= E[j-1] + goe since go <= 0)
                    //   H[j] = max(E[j], T[j])
                    // so consecutive cells of a row are linked by ONE VIADDMNMX instead of three ALU ops,
                    // at the price of one more ALU op per cell: pays when few warps share an SM.
                    int32_t tg_prev = g_in;
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        int32_t d, tg;
                        if constexpr (CODED) d = add_fma(gdiag, prmt_s((uint32_t)rb, rhi, (uint32_t)acol[j]), one);
                        else d = gdiag + ((acol[j] == rb) ? sub_match : sub_mis);
                        F[j] = __viaddmax_s32(F[j], ext, Gp[j]);
                        if constexpr (CODED) tg = add_fma(__vimax_s32_relu(F[j], d), goe, one);
                        else tg = __vimax_s32_relu(F[j], d) + goe;                 // T[j] + goe
                        e = __viaddmax_s32(e, ext, tg_prev);                       // E[i][j]
                        gdiag = Gp[j];
                        gleft = __viaddmax_s32(e, goe, tg);                        // H[i][j] + goe
                        Gp[j] = gleft;
                        tg_prev = tg;
                        if (j & 1) bestg = __vimax3_s32(bestg, Gp[j - 1], gleft);
                        else if (j == K - 1) bestg = max(bestg, gleft);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        int32_t d;
                        if constexpr (CODED) d = add_fma(gdiag, prmt_s((uint32_t)rb, rhi, (uint32_t)acol[j]), one);
                        else d = gdiag + ((acol[j] == rb) ? sub_match : sub_mis);
                        e = __viaddmax_s32(e, ext, gleft);
                        F[j] = __viaddmax_s32(F[j], ext, Gp[j]);
                        const int32_t hcell = __vimax3_s32_relu(e, F[j], d);
                        gdiag = Gp[j];
                        if constexpr (CODED) gleft = add_fma(hcell, goe, one); else gleft = hcell + goe;
                        Gp[j] = gleft;
                        if (j & 1) bestg = __vimax3_s32(bestg, Gp[j - 1], gleft);
                        else if (j == K - 1) bestg = max(bestg, gleft);
                    }
                }
                g_out = gleft;
                e_out = e;
                // lane 31 has just finished row s - 31 of the stripe's last column: stage it for the flush
                if (lane == 31) stage[wib][u] = make_int2(g_out, e_out);
            }
            {
                // inputs of the next block
                const int r = s0 + 32 + lane;
                nb = 0x200;
                nx = make_int4(goe, gst - 1, goe, gst - 1);
                if constexpr (CODED) { uint32_t lo; row_table(r, lo, nhi); nb = (int32_t)lo; }
                if (r < lb) {
                    if constexpr (!CODED) nb = g.b[r];
                    if (!left_edge) nx = ld_entry(g.bnd + r, in_remote);
                }
            }
            // hand the rows completed in this block (s0 - 31 .. s0) on with one 128-bit store per lane
            if (out_bnd != nullptr) {
                __syncwarp();
                const int r = s0 - 31 + lane;
                if (lane < send && r >= 0 && r < lb) {
                    const int2 ge = stage[wib][lane];
                    st_entry(out_bnd + r, make_int4(ge.x, gst, ge.y, gst), out_remote);
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) bestg = max(bestg, __shfl_xor_sync(0xffffffffu, bestg, m));
    const int32_t best = bestg - goe;
    if (lane == 0 && best > 0) atomicMax(g.best, best);
}

// Two rows per step (symbol-coded cells only).  With one warp per scheduler a row step is a latency chain --
// shuffle in, K dependent cells, shuffle out: ~200 clocks whatever K is -- so lane t advances TWO rows per
// step (rows 2(s-t), 2(s-t)+1): the shuffle latency and the loop are paid once per two rows, and the two
// rows' chains run one column apart, which doubles the instruction-level parallelism a lone warp offers.
// A block is 32 steps = 64 rows; everything else (tagged entries, staged flush, uniform polling) as above.
template <int K, bool SHORT>
__global__ void __launch_bounds__(LONG_WARPS * 32)
sw_long2_kernel(LongArgs g)
{
    constexpr int W = 32 * K;
    constexpr int RING = 128;
    __shared__ int32_t r_lo[LONG_WARPS][RING];
    __shared__ int32_t r_hi[LONG_WARPS][RING];
    __shared__ int32_t r_g[LONG_WARPS][RING];
    __shared__ int32_t r_e[LONG_WARPS][RING];
    __shared__ int2 stage[LONG_WARPS][64];

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * LONG_WARPS + wib;
    const int n_warps = gridDim.x * LONG_WARPS;
    const int32_t goe = g.sc.gap_open + g.sc.gap_extend;
    const int32_t ext = g.sc.gap_extend;
    const int32_t sub_match = g.sc.match - goe, sub_mis = g.sc.mismatch - goe;
    const int32_t lb = g.lb;
    const int n_stripes = (g.la + W - 1) / W;
    int32_t bestg = goe;
    const int32_t one = g.one;
    const uint32_t xb4 = (uint32_t)(uint8_t)(int8_t)sub_mis * 0x01010101u;
    const uint32_t mxor = (uint32_t)(uint8_t)(int8_t)sub_mis ^ (uint32_t)(uint8_t)(int8_t)sub_match;
    auto row_table = [&](int32_t r, uint32_t &lo, uint32_t &hi) {
        lo = xb4; hi = xb4;
        if (r < lb) {
            const uint32_t c = g.lut[g.b[r]];
            if (c < 4) lo ^= mxor << (8 * c); else hi ^= mxor << (8 * (c - 4));
        }
    };

    for (int st = warp; st < n_stripes; st += n_warps) {
        const int c0 = st * W + lane * K;
        int32_t acol[K], Gp[K], F[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const uint32_t c = (c0 + j < g.la) ? (uint32_t)g.lut[g.a[c0 + j]] : 7u;
            acol[j] = (int32_t)(c | ((8u | c) * 0x1110u));
            Gp[j] = goe;
            F[j] = goe;
        }
        int32_t g_out0 = goe, e_out0 = goe, g_out1 = goe, e_out1 = goe, g_in_prev = goe;
        const int32_t gst = g.stripe_base + st;
        const bool left_edge = (gst == 0);
        const bool last = (st == n_stripes - 1);
        int4 *out_bnd = last ? g.next_bnd : g.bnd;
        const bool out_remote = last;
        const bool in_remote = (st == 0);
        const int S = (lb + 1) / 2 + 31;                 // steps: lane 31 finishes row lb-1 at step (lb-1)/2 + 31

        // inputs of the first 64 rows: this lane loads rows lane and 32 + lane of every block
        const int4 fresh = make_int4(goe, gst - 1, goe, gst - 1);
        int4 nxa = fresh, nxb = fresh;
        uint32_t la_lo, la_hi, lb_lo, lb_hi;
        row_table(lane, la_lo, la_hi);
        row_table(32 + lane, lb_lo, lb_hi);
        if (!left_edge) {
            if (lane < lb) nxa = ld_entry(g.bnd + lane, in_remote);
            if (32 + lane < lb) nxb = ld_entry(g.bnd + 32 + lane, in_remote);
        }
        for (int s0 = 0; s0 < S; s0 += 32) {
            {
                const int ra = 2 * s0 + lane, rb_ = ra + 32;
                if (!left_edge) {
                    unsigned ns = 32;
                    while (__any_sync(0xffffffffu, nxa.y != gst - 1 || nxa.w != gst - 1 || nxb.y != gst - 1 || nxb.w != gst - 1)) {
                        __nanosleep(ns);
                        if (ns < 512) ns *= 2;
                        if (ra < lb && (nxa.y != gst - 1 || nxa.w != gst - 1)) nxa = ld_entry(g.bnd + ra, in_remote);
                        if (rb_ < lb && (nxb.y != gst - 1 || nxb.w != gst - 1)) nxb = ld_entry(g.bnd + rb_, in_remote);
                    }
                }
                r_lo[wib][ra & (RING - 1)] = (int32_t)la_lo;  r_hi[wib][ra & (RING - 1)] = (int32_t)la_hi;
                r_g[wib][ra & (RING - 1)] = nxa.x;            r_e[wib][ra & (RING - 1)] = nxa.z;
                r_lo[wib][rb_ & (RING - 1)] = (int32_t)lb_lo; r_hi[wib][rb_ & (RING - 1)] = (int32_t)lb_hi;
                r_g[wib][rb_ & (RING - 1)] = nxb.x;           r_e[wib][rb_ & (RING - 1)] = nxb.z;
            }
            __syncwarp();
            const int send = min(32, S - s0);
            // two steps per loop trip: 90.6 -> 83.3 ms on the 125 kbp x 1 Mbp share
#pragma unroll 2
            for (int u = 0; u < send; ++u) {
                const int s = s0 + u;
                const bool live = (s - lane) >= 0;
                const int slot0 = (2 * (s - lane)) & (RING - 1), slot1 = slot0 + 1;
                uint32_t t0lo = xb4, t0hi = xb4, t1lo = xb4, t1hi = xb4;
                if (live) {
                    t0lo = (uint32_t)r_lo[wib][slot0]; t0hi = (uint32_t)r_hi[wib][slot0];
                    t1lo = (uint32_t)r_lo[wib][slot1]; t1hi = (uint32_t)r_hi[wib][slot1];
                }
                int32_t g_in0 = __shfl_up_sync(0xffffffffu, g_out0, 1);
                int32_t e0 = __shfl_up_sync(0xffffffffu, e_out0, 1);
                int32_t g_in1 = __shfl_up_sync(0xffffffffu, g_out1, 1);
                int32_t e1 = __shfl_up_sync(0xffffffffu, e_out1, 1);
                if (lane == 0) {
                    g_in0 = r_g[wib][slot0]; e0 = r_e[wib][slot0];
                    g_in1 = r_g[wib][slot1]; e1 = r_e[wib][slot1];
                }
                int32_t gdiag0 = g_in_prev;              // (H+goe)[i0-1][c0-1]
                g_in_prev = g_in1;
                int32_t gdiag1 = g_in0;                  // (H+goe)[i0][c0-1]
                int32_t gleft0 = g_in0, gleft1 = g_in1;
                if constexpr (SHORT) {
                    int32_t tgp0 = g_in0, tgp1 = g_in1;
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        const int32_t d0 = add_fma(gdiag0, prmt_s(t0lo, t0hi, (uint32_t)acol[j]), one);
                        const int32_t F0 = __viaddmax_s32(F[j], ext, Gp[j]);
                        const int32_t tg0 = add_fma(__vimax_s32_relu(F0, d0), goe, one);
                        e0 = __viaddmax_s32(e0, ext, tgp0);
                        gdiag0 = Gp[j];
                        const int32_t g0 = __viaddmax_s32(e0, goe, tg0);          // (H+goe)[i0][j]
                        const int32_t d1 = add_fma(gdiag1, prmt_s(t1lo, t1hi, (uint32_t)acol[j]), one);
                        const int32_t F1 = __viaddmax_s32(F0, ext, g0);
                        const int32_t tg1 = add_fma(__vimax_s32_relu(F1, d1), goe, one);
                        e1 = __viaddmax_s32(e1, ext, tgp1);
                        gdiag1 = g0;
                        const int32_t g1 = __viaddmax_s32(e1, goe, tg1);          // (H+goe)[i1][j]
                        tgp0 = tg0; tgp1 = tg1;
                        gleft0 = g0; gleft1 = g1;
                        Gp[j] = g1; F[j] = F1;
                        bestg = __vimax3_s32(bestg, g0, g1);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        const int32_t d0 = add_fma(gdiag0, prmt_s(t0lo, t0hi, (uint32_t)acol[j]), one);
                        e0 = __viaddmax_s32(e0, ext, gleft0);
                        const int32_t F0 = __viaddmax_s32(F[j], ext, Gp[j]);
                        const int32_t g0 = add_fma(__vimax3_s32_relu(e0, F0, d0), goe, one);
                        gdiag0 = Gp[j];
                        const int32_t d1 = add_fma(gdiag1, prmt_s(t1lo, t1hi, (uint32_t)acol[j]), one);
                        e1 = __viaddmax_s32(e1, ext, gleft1);
                        const int32_t F1 = __viaddmax_s32(F0, ext, g0);
                        const int32_t g1 = add_fma(__vimax3_s32_relu(e1, F1, d1), goe, one);
                        gdiag1 = g0;
                        gleft0 = g0; gleft1 = g1;
                        Gp[j] = g1; F[j] = F1;
                        bestg = __vimax3_s32(bestg, g0, g1);
                    }
                }
                g_out0 = gleft0; e_out0 = e0; g_out1 = gleft1; e_out1 = e1;
                if (lane == 31) {
                    stage[wib][2 * u] = make_int2(g_out0, e_out0);
                    stage[wib][2 * u + 1] = make_int2(g_out1, e_out1);
                }
            }
            // inputs of the next block
            {
                const int ra = 2 * (s0 + 32) + lane, rb_ = ra + 32;
                row_table(ra, la_lo, la_hi);
                row_table(rb_, lb_lo, lb_hi);
                nxa = fresh; nxb = fresh;
                if (!left_edge) {
                    if (ra < lb) nxa = ld_entry(g.bnd + ra, in_remote);
                    if (rb_ < lb) nxb = ld_entry(g.bnd + rb_, in_remote);
                }
            }
            // hand on the rows lane 31 finished in this block: rows 2(s0-31) .. 2(s0-31) + 2*send - 1
            if (out_bnd != nullptr) {
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int q = lane + 32 * h;
                    const int r = 2 * (s0 - 31) + q;
                    if (q < 2 * send && r >= 0 && r < lb) {
                        const int2 ge = stage[wib][q];
                        st_entry(out_bnd + r, make_int4(ge.x, gst, ge.y, gst), out_remote);
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) bestg = max(bestg, __shfl_xor_sync(0xffffffffu, bestg, m));
    const int32_t best = bestg - goe;
    if (lane == 0 && best > 0) atomicMax(g.best, best);
}

template <int K, bool SHORT> int long_launch2(const LongArgs &args, int n_stripes, cudaStream_t st)
{
    int dev = 0, sms = 0, per_sm = 0;
    AGX_CUDA(cudaGetDevice(&dev));
    AGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sw_long2_kernel<K, SHORT>, LONG_WARPS * 32, 0));
    if (per_sm < 1) return fail(AGX_ECUDA, "sw_long: kernel does not fit on an SM");
    int blocks = sms * per_sm;
    const int want = (n_stripes + LONG_WARPS - 1) / LONG_WARPS;
    if (blocks > want) blocks = want;
    LongArgs a = args;
    void *params[] = {&a};
    AGX_CUDA(cudaLaunchCooperativeKernel((const void *)sw_long2_kernel<K, SHORT>, dim3(blocks), dim3(LONG_WARPS * 32),
                                         params, 0, st));
    count_launch();
    return AGX_OK;
}

template <int K, bool SHORT, bool CODED> int long_launch(const LongArgs &args, int n_stripes, cudaStream_t st)
{
    int dev = 0, sms = 0, per_sm = 0;
    AGX_CUDA(cudaGetDevice(&dev));
    AGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sw_long_kernel<K, SHORT, CODED>, LONG_WARPS * 32, 0));
    if (per_sm < 1) return fail(AGX_ECUDA, "sw_long: kernel does not fit on an SM");
    int blocks = sms * per_sm;                        // all co-resident: required by the stripe wavefront
    const int want = (n_stripes + LONG_WARPS - 1) / LONG_WARPS;
    if (blocks > want) blocks = want;
    LongArgs a = args;
    void *params[] = {&a};
    AGX_CUDA(cudaLaunchCooperativeKernel((const void *)sw_long_kernel<K, SHORT, CODED>, dim3(blocks), dim3(LONG_WARPS * 32),
                                         params, 0, st));
    count_launch();
    return AGX_OK;
}

int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return (e && atoi(e) > 0) ? atoi(e) : dflt;
}

// Stripe width.  The stripe wavefront costs (stripes x hop + rows) row steps, and a step costs about
// warps-per-scheduler x (10 K + 25) issue slots: wide stripes shorten the first term, but the stripes must
// also fill the SM sub-partitions evenly (977 stripes on 592 schedulers run at the pace of the ones that
// hold two).  pick_k() evaluates that model over the instantiated widths.
constexpr int LONG_KS[] = {2, 4, 6, 7, 8, 10, 12, 14, 16, 20, 24, 27, 32};     // CODED kernels
constexpr int LONG_KS_RAW[] = {4, 8, 16, 32};                                  // raw-byte kernels (> 7 distinct symbols)

int pick_k(int64_t cols_per_gpu, int64_t rows, int sms, bool coded)
{
    const int forced = env_int("AGX_LONG_K", 0);
    const int *ks = coded ? LONG_KS : LONG_KS_RAW;
    const int nk = coded ? (int)(sizeof(LONG_KS) / sizeof(int)) : (int)(sizeof(LONG_KS_RAW) / sizeof(int));
    for (int i = 0; i < nk; ++i)
        if (ks[i] == forced) return forced;
    const int64_t slots = (int64_t)sms * 4;
    const double per_col = coded ? 7.0 : 10.0;                        // issue slots per cell
    double best = 0;
    int best_k = 8;
    for (int i = 0; i < nk; ++i) {
        const int k = ks[i];
        const int64_t stripes = (cols_per_gpu + 32 * k - 1) / (32 * k);
        const int64_t w = (stripes + slots - 1) / slots;
        const int regs = 40 + 4 * k;                                  // 3 K state + temporaries
        if (w * 4 * 32 * regs > 65536) continue;                      // would not be co-resident
        const double cost = ((double)stripes * 64 + (double)rows) * (double)w * (per_col * k + 25.0);
        if (best == 0 || cost < best) { best = cost; best_k = k; }
    }
    return best_k;
}

template <int I = 0> int long_dispatch_coded(int k, bool short_chain, bool two_rows, const LongArgs &args, int n, cudaStream_t st)
{
    if constexpr (I < (int)(sizeof(LONG_KS) / sizeof(LONG_KS[0]))) {
        if (LONG_KS[I] == k) {
            if (two_rows)
                return short_chain ? long_launch2<LONG_KS[I], true>(args, n, st) : long_launch2<LONG_KS[I], false>(args, n, st);
            return short_chain ? long_launch<LONG_KS[I], true, true>(args, n, st)
                               : long_launch<LONG_KS[I], false, true>(args, n, st);
        }
        return long_dispatch_coded<I + 1>(k, short_chain, two_rows, args, n, st);
    } else {
        return fail(AGX_EINVAL, "sw_long: stripe width not instantiated");
    }
}
template <int I = 0> int long_dispatch_raw(int k, bool short_chain, const LongArgs &args, int n, cudaStream_t st)
{
    if constexpr (I < (int)(sizeof(LONG_KS_RAW) / sizeof(LONG_KS_RAW[0]))) {
        if (LONG_KS_RAW[I] == k)
            return short_chain ? long_launch<LONG_KS_RAW[I], true, false>(args, n, st)
                               : long_launch<LONG_KS_RAW[I], false, false>(args, n, st);
        return long_dispatch_raw<I + 1>(k, short_chain, args, n, st);
    } else {
        return fail(AGX_EINVAL, "sw_long: stripe width not instantiated");
    }
}

int long_dispatch(int k, const LongArgs &args, cudaStream_t st)
{
    const int n = (args.la + 32 * k - 1) / (32 * k);
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // one warp per scheduler: latency-bound, take the short E chain; more: issue-bound, take the lean one
    bool short_chain = n <= sms * 4;
    if (const char *e = getenv("AGX_LONG_CHAIN")) short_chain = atoi(e) != 0;
    // one warp per scheduler: two rows per step (see sw_long2_kernel)
    bool two_rows = n <= sms * 4;
    if (const char *e = getenv("AGX_LONG_ROWS")) two_rows = atoi(e) == 2;
    return args.lut ? long_dispatch_coded<0>(k, short_chain, two_rows, args, n, st) : long_dispatch_raw<0>(k, short_chain, args, n, st);
}

// byte -> symbol code for sequences with at most 7 distinct bytes (code 7 is the "matches nothing" padding);
// false when there are more, or when the scores do not fit the byte table of the CODED kernels
bool build_lut(const uint32_t present[8], SwScoring sc, uint8_t lut[256])
{
    if (getenv("AGX_LONG_RAW")) return false;                         // A/B switch: raw-byte kernels only
    const int32_t goe = sc.gap_open + sc.gap_extend;
    if (sc.match - goe > 127 || sc.mismatch - goe < -128 || sc.match - goe < -128 || sc.mismatch - goe > 127) return false;
    int n = 0;
    for (int c = 0; c < 256; ++c) {
        lut[c] = 7;
        if (present[c >> 5] >> (c & 31) & 1u) {
            if (n == 7) return false;
            lut[c] = (uint8_t)n++;
        }
    }
    return true;
}

__global__ void __launch_bounds__(256)
byte_presence_kernel(const uint8_t *__restrict__ a, int64_t la, const uint8_t *__restrict__ b, int64_t lb,
                     uint32_t *__restrict__ present)
{
    __shared__ uint32_t s_mask[8];
    if (threadIdx.x < 8) s_mask[threadIdx.x] = 0;
    __syncthreads();
    uint32_t m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < la + lb; i += stride) {
        const uint32_t c = i < la ? a[i] : b[i - la];
#pragma unroll
        for (int w = 0; w < 8; ++w) m[w] |= ((c >> 5) == (uint32_t)w) ? (1u << (c & 31)) : 0u;
    }
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        uint32_t v = m[w];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicOr(&s_mask[w], v);
    }
    __syncthreads();
    if (threadIdx.x < 8 && s_mask[threadIdx.x]) atomicOr(&present[threadIdx.x], s_mask[threadIdx.x]);
}

}  // namespace

// Device-resident single-GPU form: a (columns) and b (rows) are device pointers; *d_best receives the
// score.  Work is enqueued on st; scratch comes from ws.
int sw_long_device(SwLongWorkspace &ws, const uint8_t *d_a, int64_t la, const uint8_t *d_b, int64_t lb,
                   SwScoring sc, int32_t *d_best, cudaStream_t st)
{
    if (la <= 0 || lb <= 0) return fail(AGX_EINVAL, "sw_long: empty sequence");
    if (la > INT32_MAX - 1024 || lb > INT32_MAX / 2 - 1024) return fail(AGX_ERANGE, "sw_long: sequence too long");
    int dev = 0, sms = 148;
    AGX_CUDA(cudaGetDevice(&dev));
    AGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t need = 4 * lb + 4 + 8 + 64;              // int32 words: entries, best, presence mask, code table
    if (need > ws.cap) {
        if (ws.buf) cudaFree(ws.buf);
        ws.buf = nullptr; ws.cap = 0;
        AGX_CUDA(cudaMalloc(&ws.buf, (size_t)need * sizeof(int32_t)));
        ws.cap = need;
    }
    // which bytes occur?  <= 7 distinct ones (DNA, with or without N / newline) take the CODED kernels
    uint32_t *d_present = reinterpret_cast<uint32_t *>(ws.buf + 4 * lb + 4);
    uint8_t *d_lut = reinterpret_cast<uint8_t *>(ws.buf + 4 * lb + 4 + 8);
    AGX_CUDA(cudaMemsetAsync(d_present, 0, 8 * sizeof(uint32_t), st));
    byte_presence_kernel<<<sms * 4, 256, 0, st>>>(d_a, la, d_b, lb, d_present);
    count_launch();
    uint32_t present[8];
    AGX_CUDA(cudaMemcpyAsync(present, d_present, sizeof present, cudaMemcpyDeviceToHost, st));
    AGX_CUDA(cudaStreamSynchronize(st));
    uint8_t lut[256];
    const bool coded = build_lut(present, sc, lut);
    if (coded) AGX_CUDA(cudaMemcpyAsync(d_lut, lut, sizeof lut, cudaMemcpyHostToDevice, st));
    const int k = pick_k(la, lb, sms, coded);
    LongArgs args;
    args.lut = coded ? d_lut : nullptr;
    args.one = 1;
    args.a = d_a; args.la = (int32_t)la; args.b = d_b; args.lb = (int32_t)lb;
    args.bnd = reinterpret_cast<int4 *>(ws.buf);
    args.next_bnd = nullptr;
    args.best = d_best;
    args.stripe_base = 0;
    args.sc = sc;
    AGX_CUDA(cudaMemsetAsync(args.bnd, 0xff, (size_t)lb * sizeof(int4), st));     // tag -1: written by nobody
    AGX_CUDA(cudaMemsetAsync(d_best, 0, sizeof(int32_t), st));
    return long_dispatch(k, args, st);
}

void sw_long_workspace_free(SwLongWorkspace &ws)
{
    if (ws.buf) cudaFree(ws.buf);
    if (ws.seq) cudaFree(ws.seq);
    ws = SwLongWorkspace();
}

// Host form over n_dev GPUs (device ordinals dev[], streams st[], workspaces ws[]): columns of `a`
// are split into n_dev contiguous ranges; `b` is replicated.  Returns the score in *score_out.
int sw_long_host_multi(int n_dev, const int *dev, cudaStream_t *st, SwLongWorkspace **ws, const uint8_t *a,
                       int64_t la, const uint8_t *b, int64_t lb, SwScoring sc, int32_t *score_out)
{
    if (la <= 0 || lb <= 0) { *score_out = 0; return AGX_OK; }
    if (la > INT32_MAX - 1024 || lb > INT32_MAX / 2 - 1024) return fail(AGX_ERANGE, "sw_long: sequence too long");
    if (n_dev > 1 && la < (int64_t)n_dev * 8192) n_dev = 1;       // not worth splitting
    // peer access between neighbours
    for (int gidx = 0; gidx + 1 < n_dev; ++gidx) {
        int can = 0;
        AGX_CUDA(cudaDeviceCanAccessPeer(&can, dev[gidx], dev[gidx + 1]));
        if (!can) { n_dev = 1; break; }
        AGX_CUDA(cudaSetDevice(dev[gidx]));
        cudaError_t e = cudaDeviceEnablePeerAccess(dev[gidx + 1], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return fail(AGX_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
    }
    std::vector<LongArgs> args(n_dev);
    std::vector<int> ks(n_dev);
    std::vector<int64_t> c_lo(n_dev + 1);
    // one stripe width everywhere, interior cuts on multiples of it: only the very last
    // stripe of the matrix may carry padding columns, whose boundary nobody consumes
    int sms0 = 148;
    AGX_CUDA(cudaDeviceGetAttribute(&sms0, cudaDevAttrMultiProcessorCount, dev[0]));
    uint32_t present[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = 0; i < la; ++i) present[a[i] >> 5] |= 1u << (a[i] & 31);
    for (int64_t i = 0; i < lb; ++i) present[b[i] >> 5] |= 1u << (b[i] & 31);
    uint8_t lut[256];
    const bool coded = build_lut(present, sc, lut);
    const int k_all = pick_k((la + n_dev - 1) / n_dev, lb, sms0, coded);
    const int64_t wcols = 32 * (int64_t)k_all;
    for (int gidx = 0; gidx <= n_dev; ++gidx)
        c_lo[gidx] = (gidx == n_dev) ? la : std::min<int64_t>(la, (la * gidx / n_dev + wcols / 2) / wcols * wcols);
    // allocate + upload on every GPU, clear the tags, then make sure ALL GPUs are clear before any launch
    int32_t stripe_base = 0;
    for (int gidx = 0; gidx < n_dev; ++gidx) {
        AGX_CUDA(cudaSetDevice(dev[gidx]));
        int sms = 148;
        AGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev[gidx]));
        const int64_t cols = c_lo[gidx + 1] - c_lo[gidx];
        const int k = k_all;
        ks[gidx] = k;
        const int64_t n_stripes = (cols + 32 * k - 1) / (32 * k);
        SwLongWorkspace &w = *ws[gidx];
        const int64_t need = 4 * lb + 4;                   // one 16-byte entry per row + the best score
        if (need > w.cap) {
            if (w.buf) cudaFree(w.buf);
            w.buf = nullptr; w.cap = 0;
            AGX_CUDA(cudaMalloc(&w.buf, (size_t)need * sizeof(int32_t)));
            w.cap = need;
        }
        const int64_t seq_need = cols + lb + 64 + 256;     // + the byte -> code table
        if (seq_need > w.cap_seq) {
            if (w.seq) cudaFree(w.seq);
            w.seq = nullptr; w.cap_seq = 0;
            AGX_CUDA(cudaMalloc(&w.seq, (size_t)seq_need));
            w.cap_seq = seq_need;
        }
        AGX_CUDA(cudaMemcpyAsync(w.seq, a + c_lo[gidx], (size_t)cols, cudaMemcpyHostToDevice, st[gidx]));
        AGX_CUDA(cudaMemcpyAsync(w.seq + cols, b, (size_t)lb, cudaMemcpyHostToDevice, st[gidx]));
        uint8_t *d_lut = w.seq + (cols + lb + 63) / 64 * 64;
        if (coded) AGX_CUDA(cudaMemcpyAsync(d_lut, lut, sizeof lut, cudaMemcpyHostToDevice, st[gidx]));
        LongArgs &x = args[gidx];
        x.lut = coded ? d_lut : nullptr;
        x.one = 1;
        x.a = w.seq; x.la = (int32_t)cols; x.b = w.seq + cols; x.lb = (int32_t)lb;
        x.bnd = reinterpret_cast<int4 *>(w.buf);
        x.best = w.buf + 4 * lb;
        x.next_bnd = nullptr;
        x.stripe_base = stripe_base;
        stripe_base += (int32_t)n_stripes;
        x.sc = sc;
        AGX_CUDA(cudaMemsetAsync(x.bnd, 0xff, (size_t)lb * sizeof(int4), st[gidx]));   // tag -1: written by nobody
        AGX_CUDA(cudaMemsetAsync(x.best, 0, sizeof(int32_t), st[gidx]));
    }
    for (int gidx = 0; gidx + 1 < n_dev; ++gidx) args[gidx].next_bnd = args[gidx + 1].bnd;
    for (int gidx = 0; gidx < n_dev; ++gidx) {
        AGX_CUDA(cudaSetDevice(dev[gidx]));
        AGX_CUDA(cudaStreamSynchronize(st[gidx]));
    }
    for (int gidx = 0; gidx < n_dev; ++gidx) {
        AGX_CUDA(cudaSetDevice(dev[gidx]));
        int rc = long_dispatch(ks[gidx], args[gidx], st[gidx]);
        if (rc != AGX_OK) return rc;
    }
    int32_t best = 0;
    for (int gidx = 0; gidx < n_dev; ++gidx) {
        AGX_CUDA(cudaSetDevice(dev[gidx]));
        int32_t v = 0;
        AGX_CUDA(cudaMemcpyAsync(&v, args[gidx].best, sizeof v, cudaMemcpyDeviceToHost, st[gidx]));
        AGX_CUDA(cudaStreamSynchronize(st[gidx]));
        best = std::max(best, v);
    }
    *score_out = best;
    return AGX_OK;
}

}  // namespace agx
