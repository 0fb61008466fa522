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
    const int2 *rowtab;       // sw_longr_kernel: 8-byte score table of every row (long_rowtab_kernel), padded
    int32_t bsteps;           // sw_longr_kernel: row steps per hand-off block (<= LR_BMAX)
    int32_t slack;            // sw_longr_kernel: blocks a stripe lets its left neighbour get ahead before it starts
    int32_t req_eighths;      // sw_longr_kernel: the next block's entries are requested at step send * req_eighths / 8
    // END CELL (sw_longr_kernel<..., ENDS>): the cell the reference's running maximum comes from, as a 64-bit key
    //     H << 43 | (2^22 - 1 - (row + column)) << 21 | (2^21 - 1 - ix)
    // ix = the index along the reference's sx (the shorter line, line 1 on ties): the column when col_is_sx,
    // else the row.  atomicMax over all warps (and, on the host, over the GPUs).
    unsigned long long *best_key;
    int32_t col_base;         // global column of this GPU's column 0
    int32_t col_is_sx;
    int32_t k32;              // 32, opaque to ptxas (key = (H + goe) * 32 + tag as one IMAD)
};
constexpr uint32_t LONG_DMAX = (1u << 22) - 1u, LONG_XMAX = (1u << 21) - 1u;

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
// row steps of sw_longr_kernel per loop trip
#ifndef AGX_LONGR_UNROLL
#define AGX_LONGR_UNROLL 1
#endif
constexpr int LONGR_UNROLL = AGX_LONGR_UNROLL;

// A boundary entry may be (re)written by another SM or by a peer GPU while it is polled: relaxed (strong)
// accesses at the narrowest scope that covers writer and reader -- .gpu inside one GPU, .sys across NVLink.
// (ld/st.volatile and plain .cg stores were measured too: no difference.)  Requesting the next block's entries
// a block ahead was also tried: it needs a two-block start-up slack and came out 5-10 % slower.
// An entry is two 64-bit halves, {H+goe, tag} and {E, tag}, read and written as .v2.b64: the memory model makes each
// 64-bit ELEMENT of a vector access single-copy atomic, so a value can never be seen with another write's tag
// (a .v4.s32 access only guarantees that per 32-bit element), also for peer stores over NVLink.
// What a load of an entry returns: the two 64-bit halves as they came.  The fields are taken apart where they are
// USED -- unpacking them next to the load made the warp wait for every early request on the spot (15 % of a lone
// warp's stall samples sat on those moves: profiles/r2y_longr_7_4_ncu.txt).
struct LEntry {
    unsigned long long a, b;
    __device__ __forceinline__ int32_t x() const { return (int32_t)(uint32_t)a; }            // H + goe
    __device__ __forceinline__ int32_t y() const { return (int32_t)(uint32_t)(a >> 32); }    // tag
    __device__ __forceinline__ int32_t z() const { return (int32_t)(uint32_t)b; }            // E
    __device__ __forceinline__ int32_t w() const { return (int32_t)(uint32_t)(b >> 32); }    // tag
};
__device__ __forceinline__ LEntry make_lentry(int32_t x, int32_t y, int32_t z, int32_t w)
{
    LEntry e;
    e.a = (unsigned long long)(uint32_t)x | ((unsigned long long)(uint32_t)y << 32);
    e.b = (unsigned long long)(uint32_t)z | ((unsigned long long)(uint32_t)w << 32);
    return e;
}
__device__ __forceinline__ LEntry ld_entry(const int4 *p, bool sys)
{
    LEntry e;
    if (sys)
        asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(e.a), "=l"(e.b) : "l"(p) : "memory");
    else
        asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(e.a), "=l"(e.b) : "l"(p) : "memory");
    return e;
}
__device__ __forceinline__ void st_entry(int4 *p, int4 v, bool sys)
{
    const unsigned long long a = (unsigned long long)(uint32_t)v.x | ((unsigned long long)(uint32_t)v.y << 32);
    const unsigned long long b = (unsigned long long)(uint32_t)v.z | ((unsigned long long)(uint32_t)v.w << 32);
    if (sys)
        asm volatile("st.relaxed.sys.global.v2.b64 [%0], {%1, %2};" :: "l"(p), "l"(a), "l"(b) : "memory");
    else
        asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}

// Raw-byte kernel: sequences with more than 7 distinct bytes (or scores that do not fit a byte table).  One row per
// step, substitution score by ISETP + SEL: 7.5 (lean chain) / 8.5 (short chain) ALU-pipe instructions per cell.
// Sequences with at most 7 distinct bytes -- all DNA -- take the symbol-coded sw_longr_kernel below.
template <int K, bool SHORT>
__global__ void __launch_bounds__(LONG_WARPS * 32)
sw_long_kernel(LongArgs g)
{
    constexpr int W = 32 * K;
    __shared__ int32_t r_byte[LONG_WARPS][LONG_RING];      // row byte
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

    for (int st = warp; st < n_stripes; st += n_warps) {
        const int c0 = st * W + lane * K;
        int32_t acol[K], Gp[K], F[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            acol[j] = (c0 + j < g.la) ? (int32_t)g.a[c0 + j] : 0x100;        // 0x100 never equals a byte
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
        LEntry nx = make_lentry(goe, gst - 1, goe, gst - 1);
        if (lane < lb) {
            nb = g.b[lane];
            if (!left_edge) nx = ld_entry(g.bnd + lane, in_remote);
        }
        for (int s0 = 0; s0 < S; s0 += 32) {
            {
                const int r = s0 + lane;
                if (!left_edge) {
                    // wait until the left neighbour has handed these 32 rows on
                    unsigned ns = 32;
                    while (__any_sync(0xffffffffu, nx.y() != gst - 1 || nx.w() != gst - 1)) {
                        __nanosleep(ns);
                        if (ns < 512) ns *= 2;
                        if (r < lb && (nx.y() != gst - 1 || nx.w() != gst - 1)) nx = ld_entry(g.bnd + r, in_remote);
                    }
                }
                r_byte[wib][r & (LONG_RING - 1)] = nb;
                r_g[wib][r & (LONG_RING - 1)] = nx.x();
                r_e[wib][r & (LONG_RING - 1)] = nx.z();
            }
            __syncwarp();
            const int send = min(32, S - s0);
#pragma unroll LONG_UNROLL
            for (int u = 0; u < send; ++u) {
                const int s = s0 + u;
                const int slot = (s - lane) & (LONG_RING - 1);
                const int32_t rb = (s - lane >= 0) ? r_byte[wib][slot] : 0x200;
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
                        const int32_t d = gdiag + ((acol[j] == rb) ? sub_match : sub_mis);
                        F[j] = __viaddmax_s32(F[j], ext, Gp[j]);
                        const int32_t tg = __vimax_s32_relu(F[j], d) + goe;        // T[j] + goe
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
                        const int32_t d = gdiag + ((acol[j] == rb) ? sub_match : sub_mis);
                        e = __viaddmax_s32(e, ext, gleft);
                        F[j] = __viaddmax_s32(F[j], ext, Gp[j]);
                        const int32_t hcell = __vimax3_s32_relu(e, F[j], d);
                        gdiag = Gp[j];
                        gleft = hcell + goe;
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
                nx = make_lentry(goe, gst - 1, goe, gst - 1);
                if (r < lb) {
                    nb = g.b[r];
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

// ------------------------------------------------------------------------------------------------------
// R rows per step.  The general form of the two kernels above for symbol-coded sequences: lane t holds K
// columns and advances R rows per step (rows R(s-t) .. R(s-t)+R-1), i.e. an R x K tile of cells whose R row
// chains run one column apart.  Why: with one warp per scheduler (the per-GPU share of an 8-GPU run) a step
// is a latency chain -- shuffle in, dependent cells, shuffle out -- and the per-step costs (2R shuffles, row
// tables, lane-0 boundary, staging, loop) are paid once per R*K cells; the total run is
//     (stripes * (32 + B) + rows / R) steps            (32 = lane skew, B = steps per hand-off block)
// so R trades the row term against the stripe-fill term, and B (a run-time argument) trims the fill.
//   - Row score tables come from a pre-pass (long_rowtab_kernel: one coalesced 8-byte load per row instead of
//     two dependent ones) and are requested a block ahead.
//   - The ring holds neutral rows for the 31 steps in which high lanes have no row yet, so the step has no
//     "is this lane live" predicate.
//   - DP4A: ONE PRMT fetches the substitution bytes of four columns and each cell adds its byte to the
//     diagonal with dp4a (one-hot multiplier) -- PRMT work per cell drops from 1 to 1/4 ALU-pipe instruction.
// Everything else (tagged 16-byte entries, staged flush, warp-uniform polling, peer stores) as above.
constexpr int LR_BMAX = 32;

template <int R> __device__ __forceinline__ void lds_vec(const int32_t *p, int32_t (&v)[R])
{
    if constexpr (R == 4) { const int4 x = *reinterpret_cast<const int4 *>(p); v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; }
    else if constexpr (R % 2 == 0) {
#pragma unroll
        for (int i = 0; i < R; i += 2) { const int2 x = *reinterpret_cast<const int2 *>(p + i); v[i] = x.x; v[i + 1] = x.y; }
    } else {
#pragma unroll
        for (int i = 0; i < R; ++i) v[i] = p[i];
    }
}

template <int K, int R, bool SHORT, bool DP4A, bool ENDS>
__global__ void __launch_bounds__(LONG_WARPS * 32)
sw_longr_kernel(LongArgs g)
{
    constexpr int W = 32 * K;
    constexpr int RROWS = 64 * R;                      // ring rows: 31 steps of lane skew + a block of <= 32 steps
    constexpr int K4 = (K + 3) / 4;
    __shared__ __align__(16) int32_t r_lo[LONG_WARPS][RROWS];
    __shared__ __align__(16) int32_t r_hi[LONG_WARPS][RROWS];
    __shared__ __align__(16) int32_t r_g[LONG_WARPS][RROWS];
    __shared__ __align__(16) int32_t r_e[LONG_WARPS][RROWS];
    __shared__ __align__(16) int2 stage[LONG_WARPS][LR_BMAX * R];

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * LONG_WARPS + wib;
    const int n_warps = gridDim.x * LONG_WARPS;
    const int32_t goe = g.sc.gap_open + g.sc.gap_extend;
    const int32_t ext = g.sc.gap_extend;
    const int32_t lb = g.lb;
    const int n_stripes = (g.la + W - 1) / W;
    const int B = g.bsteps;
    int32_t bestg[(R + 1) / 2];
#pragma unroll
    for (int i = 0; i < (R + 1) / 2; ++i) bestg[i] = goe;
    unsigned long long best64 = 0ull;                  // ENDS
    const int32_t one = g.one;
    const int32_t xb4 = (int32_t)((uint32_t)(uint8_t)(int8_t)(g.sc.mismatch - goe) * 0x01010101u);
    const int S = (lb + R - 1) / R + 31;               // lane 31 finishes the last row block at step (lb-1)/R + 31

    for (int st = warp; st < n_stripes; st += n_warps) {
        const int c0 = st * W + lane * K;
        // column symbols: one PRMT selector per column, or (DP4A) four 3-bit codes per selector
        int32_t acol[DP4A ? K4 : K], Gp[K], F[K];
#pragma unroll
        for (int j = 0; j < K; ++j) { Gp[j] = goe; F[j] = goe; }
        if constexpr (DP4A) {
#pragma unroll
            for (int q = 0; q < K4; ++q) {
                uint32_t sel = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = 4 * q + e;
                    const uint32_t c = (j < K && c0 + j < g.la) ? (uint32_t)g.lut[g.a[c0 + j]] : 7u;
                    sel |= c << (4 * e);
                }
                acol[q] = (int32_t)sel;
            }
        } else {
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const uint32_t c = (c0 + j < g.la) ? (uint32_t)g.lut[g.a[c0 + j]] : 7u;
                acol[j] = (int32_t)(c | ((8u | c) * 0x1110u));              // byte c, sign-extended to 32 bits
            }
        }
        int32_t g_out[R], e_out[R];
#pragma unroll
        for (int i = 0; i < R; ++i) { g_out[i] = goe; e_out[i] = goe; }
        int32_t g_in_prev = goe;
        const int32_t gst = g.stripe_base + st;                          // global stripe index = my tag
        const bool left_edge = (gst == 0);
        const bool last = (st == n_stripes - 1);
        int4 *out_bnd = last ? g.next_bnd : g.bnd;                       // nullptr: nothing to hand on
        const bool out_remote = last;
        const bool in_remote = (st == 0);
        const LEntry fresh = make_lentry(goe, gst - 1, goe, gst - 1);

        // neutral rows -31R .. -1 (what a lane sees before its first real row)
        __syncwarp();
        for (int i = lane; i < 31 * R; i += 32) {
            const int slot = (RROWS - 31 * R) + i;
            r_lo[wib][slot] = xb4; r_hi[wib][slot] = xb4; r_g[wib][slot] = goe; r_e[wib][slot] = goe;
        }
        // Start slack.  A stripe that starts the moment its first block has been handed on stays "just in time" for
        // good: every later block's entries are requested before they exist and fetched again at the top of the block
        // (an L2 round trip per block, a third of a lone warp's cycles: ncu, long_scoreboard on ld_entry).  Letting
        // the left neighbour get `slack` blocks further ahead first makes the early request find its data.
        if (!left_edge && g.slack > 0) {
            const int probe = min(lb - 1, (1 + g.slack) * B * R - 1);
            unsigned ns = 32;
            for (;;) {
                const LEntry v = ld_entry(g.bnd + probe, in_remote);
                if (v.y() == gst - 1 && v.w() == gst - 1) break;
                __nanosleep(ns);
                if (ns < 512) ns *= 2;
            }
        }
        // inputs of the first block: this lane owns rows base + 32 q + lane of every block
        int2 nt[R];
        LEntry nx[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int idx = 32 * q + lane;
            nt[q] = make_int2(xb4, xb4);
            nx[q] = fresh;
            if (idx < B * R) {
                nt[q] = g.rowtab[idx];
                if (!left_edge && idx < lb) nx[q] = ld_entry(g.bnd + idx, in_remote);
            }
        }
        for (int s0 = 0; s0 < S; s0 += B) {
            const int base = R * s0;
            if (!left_edge) {
                // wait until the left neighbour has handed this block's rows on
                unsigned ns = 20;
                for (;;) {
                    bool missing = false;
#pragma unroll
                    for (int q = 0; q < R; ++q) missing = missing || nx[q].y() != gst - 1 || nx[q].w() != gst - 1;
                    if (!__any_sync(0xffffffffu, missing)) break;
                    __nanosleep(ns);
                    if (ns < 320) ns *= 2;
#pragma unroll
                    for (int q = 0; q < R; ++q)
                        if (nx[q].y() != gst - 1 || nx[q].w() != gst - 1) nx[q] = ld_entry(g.bnd + base + 32 * q + lane, in_remote);
                }
            }
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int idx = 32 * q + lane;
                if (idx < B * R) {
                    const int slot = (base + idx) & (RROWS - 1);
                    r_lo[wib][slot] = nt[q].x; r_hi[wib][slot] = nt[q].y;
                    r_g[wib][slot] = nx[q].x();  r_e[wib][slot] = nx[q].z();
                }
            }
            __syncwarp();
            // row tables of the next block: requested now, consumed after this block's steps
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int idx = 32 * q + lane;
                nx[q] = fresh;
                if (idx < B * R) nt[q] = g.rowtab[base + B * R + idx];       // rowtab is padded past the last block
            }
            const int send = min(B, S - s0);
            const int half = (send * g.req_eighths) >> 3;
            // row tables / lane-0 boundary of this block's first step; later steps get theirs one step ahead (a lone
            // warp has nothing else to cover the shared-memory latency with)
            int32_t tlo[R], thi[R], bg[R], be[R];
            {
                const int sl = (R * (s0 - lane)) & (RROWS - 1), sl0 = (R * s0) & (RROWS - 1);
                lds_vec<R>(&r_lo[wib][sl], tlo);
                lds_vec<R>(&r_hi[wib][sl], thi);
                lds_vec<R>(&r_g[wib][sl0], bg);
                lds_vec<R>(&r_e[wib][sl0], be);
            }
#pragma unroll LONGR_UNROLL
            for (int u = 0; u < send; ++u) {
                const int s = s0 + u;
                if (u == half && !left_edge) {
                    // boundary entries of the next block: requested half a block early, so that their way through L2
                    // (or NVLink) is covered by the remaining steps; the poll at the top repeats what was not there yet
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int r = base + B * R + 32 * q + lane;
                        if (32 * q + lane < B * R && r < lb) nx[q] = ld_entry(g.bnd + r, in_remote);
                    }
                }
                int32_t plo[R], phi[R], pg[R], pe[R];             // the next step's (dropped at the end of a block)
                {
                    const int sl = (R * (s + 1 - lane)) & (RROWS - 1), sl0 = (R * (s + 1)) & (RROWS - 1);
                    lds_vec<R>(&r_lo[wib][sl], plo);
                    lds_vec<R>(&r_hi[wib][sl], phi);
                    lds_vec<R>(&r_g[wib][sl0], pg);
                    lds_vec<R>(&r_e[wib][sl0], pe);
                }
                int32_t g_in[R], e[R];
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    g_in[i] = __shfl_up_sync(0xffffffffu, g_out[i], 1);
                    e[i] = __shfl_up_sync(0xffffffffu, e_out[i], 1);
                    if (lane == 0) { g_in[i] = bg[i]; e[i] = be[i]; }
                }
                int32_t gdiag[R], gleft[R], tgp[R];
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    gdiag[i] = i == 0 ? g_in_prev : g_in[i - 1];           // (H+goe)[row-1][c0-1]
                    gleft[i] = g_in[i];
                    tgp[i] = g_in[i];
                }
                g_in_prev = g_in[R - 1];
                uint32_t sc4[R];
                int32_t rk[ENDS ? R : 1];                          // ENDS: row maximum of (H + goe) * 32 + (31 - j)
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    if constexpr (DP4A) {
                        if ((j & 3) == 0) {
#pragma unroll
                            for (int i = 0; i < R; ++i) sc4[i] = __byte_perm((uint32_t)tlo[i], (uint32_t)thi[i], (uint32_t)acol[j >> 2]);
                        }
                    }
                    int32_t up = Gp[j], f = F[j];
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        int32_t d;
                        if constexpr (DP4A) d = __dp4a((int32_t)sc4[i], (int32_t)(1u << (8 * (j & 3))), gdiag[i]);
                        else d = add_fma(gdiag[i], prmt_s((uint32_t)tlo[i], (uint32_t)thi[i], (uint32_t)acol[j]), one);
                        f = __viaddmax_s32(f, ext, up);                              // F[i][j]
                        int32_t gnew;
                        if constexpr (SHORT) {
                            const int32_t tg = add_fma(__vimax_s32_relu(f, d), goe, one);      // T + goe, T = max(F, diag + s, 0)
                            e[i] = __viaddmax_s32(e[i], ext, tgp[i]);                // E[i][j] = max(E[i][j-1] + ext, T[i][j-1] + goe)
                            gnew = __viaddmax_s32(e[i], goe, tg);                    // H[i][j] + goe
                            tgp[i] = tg;
                        } else {
                            e[i] = __viaddmax_s32(e[i], ext, gleft[i]);              // E[i][j]
                            gnew = add_fma(__vimax3_s32_relu(e[i], f, d), goe, one); // H[i][j] + goe
                        }
                        gdiag[i] = up;
                        up = gnew;
                        gleft[i] = gnew;
                        if constexpr (ENDS) {
                            // the row maximum of these keys names the row's best H and its first column
                            int32_t key;
                            asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(key) : "r"(gnew), "r"(g.k32), "r"(31 - j));
                            rk[i] = j == 0 ? key : max(rk[i], key);
                        } else {
                            // running maximum: one accumulator per pair of rows (independent chains)
                            if constexpr (R == 1) { if (j & 1) bestg[0] = __vimax3_s32(bestg[0], Gp[j - 1], gnew); else if (j == K - 1) bestg[0] = max(bestg[0], gnew); }
                            else if (i & 1) bestg[i >> 1] = __vimax3_s32(bestg[i >> 1], gleft[i - 1], gnew);
                            else if (i == R - 1) bestg[i >> 1] = max(bestg[i >> 1], gnew);
                        }
                    }
                    Gp[j] = up; F[j] = f;
                }
                if constexpr (ENDS) {
                    // widen each row's key to the order the reference visits cells in: larger H, then the earlier
                    // anti-diagonal, then the smaller ix.  Rows before the first / behind the last real row and padding
                    // columns hold values that only decay from real cells: never the maximum.
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const int32_t h = (rk[i] >> 5) - goe;
                        const uint32_t col = (uint32_t)(g.col_base + c0 + 31 - (rk[i] & 31));
                        const uint32_t row = (uint32_t)(R * (s - lane) + i);
                        const uint32_t x = g.col_is_sx ? col : row;
                        const unsigned long long k64 = ((unsigned long long)(uint32_t)max(h, 0) << 43) |
                                                       ((unsigned long long)((LONG_DMAX - (row + col)) & LONG_DMAX) << 21) |
                                                       (unsigned long long)((LONG_XMAX - x) & LONG_XMAX);
                        best64 = max(best64, k64);
                    }
                }
#pragma unroll
                for (int i = 0; i < R; ++i) { g_out[i] = gleft[i]; e_out[i] = e[i]; }
                // lane 31 has just finished rows R(s-31) .. +R-1 of the stripe's last column: stage them
                if (lane == 31) {
#pragma unroll
                    for (int i = 0; i < R; ++i) stage[wib][u * R + i] = make_int2(g_out[i], e_out[i]);
                }
#pragma unroll
                for (int i = 0; i < R; ++i) { tlo[i] = plo[i]; thi[i] = phi[i]; bg[i] = pg[i]; be[i] = pe[i]; }
            }
            // hand on the rows lane 31 finished in this block: rows R(s0-31) .. R(s0-31) + R*send - 1
            if (out_bnd != nullptr) {
                __syncwarp();
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int idx = 32 * q + lane;
                    const int r = R * (s0 - 31) + idx;
                    if (idx < R * send && r >= 0 && r < lb) {
                        const int2 ge = stage[wib][idx];
                        st_entry(out_bnd + r, make_int4(ge.x, gst, ge.y, gst), out_remote);
                    }
                }
            }
            __syncwarp();
        }
        __syncwarp();
    }
    if constexpr (ENDS) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) best64 = max(best64, __shfl_xor_sync(0xffffffffu, best64, m));
        if (lane == 0 && (best64 >> 43) > 0) {
            atomicMax(g.best_key, best64);
            atomicMax(g.best, (int32_t)(best64 >> 43));
        }
        return;
    }
    int32_t best_all = bestg[0];
#pragma unroll
    for (int i = 1; i < (R + 1) / 2; ++i) best_all = max(best_all, bestg[i]);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) best_all = max(best_all, __shfl_xor_sync(0xffffffffu, best_all, m));
    const int32_t best = best_all - goe;
    if (lane == 0 && best > 0) atomicMax(g.best, best);
}

// ------------------------------------------------------------------------------------------------------
// Software-pipelined form of sw_longr_kernel (lean chain).  An R x K tile has low parallelism in its first and
// last cell diagonals, and because the next step's first cell needs this step's last shuffle, a lone warp
// (issue is in order) idles through both: measured 440 clocks per step for 190 instructions at K = 7, R = 4.
// Here row i of a tile runs i columns behind row 0 and WRAPS into the next iteration: at time c of an
// iteration (c = 0 .. K-1) row i computes column c - i of the lane's current row block when c >= i, and
// column K + c - i of the previous row block otherwise.  Every time slot then holds exactly R independent
// cells, one per row, and all column indices stay compile-time constants.  A row's boundary value leaves for
// the next lane right after its last column (time i - 1 of the following iteration) and is consumed at time i:
// no shuffle waits at a step boundary any more.  Lane 31 finishes row block b in iteration b + 32.
template <int K, int R, bool DP4A>
__global__ void __launch_bounds__(LONG_WARPS * 32)
sw_longp_kernel(LongArgs g)
{
    static_assert(R >= 1 && R <= K, "rows per step must not exceed columns per lane");
    constexpr int W = 32 * K;
    constexpr int RROWS = 64 * R;                      // ring rows: 31 steps of lane skew + a block of <= 32 steps
    auto slot_of = [](int row) { return (row + RROWS) % RROWS; };       // rows >= -32 R; a mask when R is a power of two
    constexpr int K4 = (K + 3) / 4;
    __shared__ __align__(16) int32_t r_lo[LONG_WARPS][RROWS];
    __shared__ __align__(16) int32_t r_hi[LONG_WARPS][RROWS];
    __shared__ __align__(16) int32_t r_g[LONG_WARPS][RROWS];
    __shared__ __align__(16) int32_t r_e[LONG_WARPS][RROWS];
    __shared__ __align__(16) int2 stage[LONG_WARPS][RROWS];      // ring of finished boundary rows (row & (RROWS-1))

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * LONG_WARPS + wib;
    const int n_warps = gridDim.x * LONG_WARPS;
    const int32_t goe = g.sc.gap_open + g.sc.gap_extend;
    const int32_t ext = g.sc.gap_extend;
    const int32_t lb = g.lb;
    const int n_stripes = (g.la + W - 1) / W;
    const int B = g.bsteps;
    int32_t bestg[(R + 1) / 2];
#pragma unroll
    for (int i = 0; i < (R + 1) / 2; ++i) bestg[i] = goe;
    const int32_t one = g.one;
    const int32_t xb4 = (int32_t)((uint32_t)(uint8_t)(int8_t)(g.sc.mismatch - goe) * 0x01010101u);
    const int S = (lb + R - 1) / R + 32;               // lane 31 finishes the wrapped rows of the last block in iteration nb + 31

    for (int st = warp; st < n_stripes; st += n_warps) {
        const int c0 = st * W + lane * K;
        int32_t acol[DP4A ? K4 : K], Gp[K], F[K];
#pragma unroll
        for (int j = 0; j < K; ++j) { Gp[j] = goe; F[j] = goe; }
        if constexpr (DP4A) {
#pragma unroll
            for (int q = 0; q < K4; ++q) {
                uint32_t sel = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = 4 * q + e;
                    const uint32_t c = (j < K && c0 + j < g.la) ? (uint32_t)g.lut[g.a[c0 + j]] : 7u;
                    sel |= c << (4 * e);
                }
                acol[q] = (int32_t)sel;
            }
        } else {
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const uint32_t c = (c0 + j < g.la) ? (uint32_t)g.lut[g.a[c0 + j]] : 7u;
                acol[j] = (int32_t)(c | ((8u | c) * 0x1110u));
            }
        }
        // per-row state (a row that has not entered a real block yet is neutral: H = 0 everywhere)
        int32_t e[R], gleft[R], gdiag[R], tlo[R], thi[R], gs[R], es[R];
        uint32_t sc4[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            e[i] = goe; gleft[i] = goe; gdiag[i] = goe; tlo[i] = xb4; thi[i] = xb4; gs[i] = goe; es[i] = goe;
            sc4[i] = (uint32_t)xb4;
        }
        int32_t gcarry = goe;                                            // (H+goe) left of the row above the entering one
        const int32_t gst = g.stripe_base + st;
        const bool left_edge = (gst == 0);
        const bool last = (st == n_stripes - 1);
        int4 *out_bnd = last ? g.next_bnd : g.bnd;
        const bool out_remote = last;
        const bool in_remote = (st == 0);
        const LEntry fresh = make_lentry(goe, gst - 1, goe, gst - 1);

        __syncwarp();
        for (int i = lane; i < 31 * R; i += 32) {
            const int slot = (RROWS - 31 * R) + i;
            r_lo[wib][slot] = xb4; r_hi[wib][slot] = xb4; r_g[wib][slot] = goe; r_e[wib][slot] = goe;
        }
        int2 nt[R];
        LEntry nx[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int idx = 32 * q + lane;
            nt[q] = make_int2(xb4, xb4);
            nx[q] = fresh;
            if (idx < B * R) {
                nt[q] = g.rowtab[idx];
                if (!left_edge && idx < lb) nx[q] = ld_entry(g.bnd + idx, in_remote);
            }
        }
        for (int s0 = 0; s0 < S; s0 += B) {
            const int base = R * s0;
            if (!left_edge) {
                unsigned ns = 20;
                for (;;) {
                    bool missing = false;
#pragma unroll
                    for (int q = 0; q < R; ++q) missing = missing || nx[q].y() != gst - 1 || nx[q].w() != gst - 1;
                    if (!__any_sync(0xffffffffu, missing)) break;
                    __nanosleep(ns);
                    if (ns < 320) ns *= 2;
#pragma unroll
                    for (int q = 0; q < R; ++q)
                        if (nx[q].y() != gst - 1 || nx[q].w() != gst - 1) nx[q] = ld_entry(g.bnd + base + 32 * q + lane, in_remote);
                }
            }
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int idx = 32 * q + lane;
                if (idx < B * R) {
                    const int slot = slot_of(base + idx);
                    r_lo[wib][slot] = nt[q].x; r_hi[wib][slot] = nt[q].y;
                    r_g[wib][slot] = nx[q].x();  r_e[wib][slot] = nx[q].z();
                }
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int idx = 32 * q + lane;
                nx[q] = fresh;
                if (idx < B * R) nt[q] = g.rowtab[base + B * R + idx];
            }
            const int send = min(B, S - s0);
            const int half = send >> 1;
            // row tables / lane-0 boundary of the first iteration of this block; later ones are requested one
            // iteration ahead (a lone warp cannot hide the shared-memory latency otherwise)
            int32_t nlo[R], nhi[R], bg[R], be[R];
            {
                const int sl = R * ((s0 - lane) & 63), sl0 = R * (s0 & 63);
                lds_vec<R>(&r_lo[wib][sl], nlo);
                lds_vec<R>(&r_hi[wib][sl], nhi);
                lds_vec<R>(&r_g[wib][sl0], bg);
                lds_vec<R>(&r_e[wib][sl0], be);
            }
#pragma unroll 1
            for (int u = 0; u < send; ++u) {
                const int s = s0 + u;
                if (u == half && !left_edge) {
                    // boundary entries of the next block, requested half a block early so that their latency is
                    // covered by the remaining steps (the poll at the top repeats what had not arrived yet)
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int r = base + B * R + 32 * q + lane;
                        if (32 * q + lane < B * R && r < lb) nx[q] = ld_entry(g.bnd + r, in_remote);
                    }
                }
                // next iteration's tables (for lane 0 the last iteration of a block reads rows that are not in
                // the ring yet: those values are dropped, the block start above reads them again)
                int32_t plo[R], phi[R], pg[R], pe[R];
                {
                    const int sl = R * ((s + 1 - lane) & 63), sl0 = R * ((s + 1) & 63);
                    lds_vec<R>(&r_lo[wib][sl], plo);
                    lds_vec<R>(&r_hi[wib][sl], phi);
                    lds_vec<R>(&r_g[wib][sl0], pg);
                    lds_vec<R>(&r_e[wib][sl0], pe);
                }
                // lane 31 stages finished rows: row i of its current block (s - 31) or, wrapped, of the one before
                int2 *stage_cur = &stage[wib][R * ((s - 31) & 63)], *stage_prev = &stage[wib][R * ((s - 32) & 63)];
#pragma unroll
                for (int c = 0; c < K; ++c) {
                    // the R cells of this time slot, one per row, all independent: written stage by stage so that
                    // the instruction stream interleaves them
                    int32_t d[R], f[R], gnew[R];
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const int j = (c >= i) ? c - i : K + c - i;
                        if (j == 0) {
                            // row i enters the lane's current row block
                            const int32_t gin = (lane == 0) ? bg[i] : gs[i];
                            const int32_t ein = (lane == 0) ? be[i] : es[i];
                            gdiag[i] = gcarry;
                            gcarry = gin;
                            gleft[i] = gin;
                            e[i] = ein;
                            tlo[i] = nlo[i];
                            thi[i] = nhi[i];
                        }
                        if constexpr (DP4A) {
                            if ((j & 3) == 0) sc4[i] = __byte_perm((uint32_t)tlo[i], (uint32_t)thi[i], (uint32_t)acol[j >> 2]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const int j = (c >= i) ? c - i : K + c - i;
                        if constexpr (DP4A) d[i] = __dp4a((int32_t)sc4[i], (int32_t)(1u << (8 * (j & 3))), gdiag[i]);
                        else d[i] = add_fma(gdiag[i], prmt_s((uint32_t)tlo[i], (uint32_t)thi[i], (uint32_t)acol[j]), one);
                    }
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const int j = (c >= i) ? c - i : K + c - i;
                        f[i] = __viaddmax_s32(F[j], ext, Gp[j]);                     // F[row][j]
                    }
#pragma unroll
                    for (int i = 0; i < R; ++i) e[i] = __viaddmax_s32(e[i], ext, gleft[i]);     // E[row][j]
#pragma unroll
                    for (int i = 0; i < R; ++i) gnew[i] = __vimax3_s32_relu(e[i], f[i], d[i]);
#pragma unroll
                    for (int i = 0; i < R; ++i) gnew[i] = add_fma(gnew[i], goe, one);           // H[row][j] + goe
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const int j = (c >= i) ? c - i : K + c - i;
                        gdiag[i] = Gp[j];
                        Gp[j] = gnew[i];
                        F[j] = f[i];
                        gleft[i] = gnew[i];
                        if constexpr (R == 1) bestg[0] = max(bestg[0], gnew[i]);
                        else if (i & 1) bestg[i >> 1] = __vimax3_s32(bestg[i >> 1], gnew[i - 1], gnew[i]);
                        else if (i == R - 1) bestg[i >> 1] = max(bestg[i >> 1], gnew[i]);
                        if (j == K - 1) {
                            // the row leaves this lane: hand its boundary to the next lane (lane 31: stage it)
                            gs[i] = __shfl_up_sync(0xffffffffu, gnew[i], 1);
                            es[i] = __shfl_up_sync(0xffffffffu, e[i], 1);
                            if (lane == 31) (c >= i ? stage_cur : stage_prev)[i] = make_int2(gnew[i], e[i]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < R; ++i) { nlo[i] = plo[i]; nhi[i] = phi[i]; bg[i] = pg[i]; be[i] = pe[i]; }
            }
            // hand on the row blocks lane 31 completed by now: blocks s0 - 32 .. s0 + send - 33
            if (out_bnd != nullptr) {
                __syncwarp();
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int idx = 32 * q + lane;
                    const int r = R * (s0 - 32) + idx;
                    if (idx < R * send && r >= 0 && r < lb) {
                        const int2 ge = stage[wib][R * ((r / R) & 63) + r % R];
                        st_entry(out_bnd + r, make_int4(ge.x, gst, ge.y, gst), out_remote);
                    }
                }
            }
            __syncwarp();
        }
        __syncwarp();
    }
    int32_t best_all = bestg[0];
#pragma unroll
    for (int i = 1; i < (R + 1) / 2; ++i) best_all = max(best_all, bestg[i]);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) best_all = max(best_all, __shfl_xor_sync(0xffffffffu, best_all, m));
    const int32_t best = best_all - goe;
    if (lane == 0 && best > 0) atomicMax(g.best, best);
}

// row r -> its 8-byte score table (byte k = substitution score - goe against symbol code k; code 7 and the
// padding rows past the last one never match), laid out for one coalesced 8-byte load per row
__global__ void __launch_bounds__(256)
long_rowtab_kernel(const uint8_t *__restrict__ b, int32_t lb, int64_t n_padded, const uint8_t *__restrict__ lut,
                   SwScoring sc, int2 *__restrict__ rowtab)
{
    const int32_t goe = sc.gap_open + sc.gap_extend;
    const uint32_t xb4 = (uint32_t)(uint8_t)(int8_t)(sc.mismatch - goe) * 0x01010101u;
    const uint32_t mxor = (uint32_t)(uint8_t)(int8_t)(sc.mismatch - goe) ^ (uint32_t)(uint8_t)(int8_t)(sc.match - goe);
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_padded; r += (int64_t)gridDim.x * blockDim.x) {
        uint32_t lo = xb4, hi = xb4;
        if (r < lb) {
            const uint32_t c = lut[b[r]];
            if (c < 4) lo ^= mxor << (8 * c); else hi ^= mxor << (8 * (c - 4));
        }
        rowtab[r] = make_int2((int32_t)lo, (int32_t)hi);
    }
}
// rows of padding behind the last real row: one block of row tables is always requested ahead
constexpr int64_t LR_ROW_PAD = 512;        // > 95 R for R <= 4
__host__ __device__ constexpr int64_t lr_rows_padded(int64_t lb) { return lb + LR_ROW_PAD; }

template <int K, int R, bool SHORT, bool DP4A, bool ENDS = false> int long_launch_r(const LongArgs &args, int n_stripes, cudaStream_t st)
{
    int dev = 0, sms = 0, per_sm = 0;
    AGX_CUDA(cudaGetDevice(&dev));
    AGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sw_longr_kernel<K, R, SHORT, DP4A, ENDS>, LONG_WARPS * 32, 0));
    if (per_sm < 1) return fail(AGX_ECUDA, "sw_long: kernel does not fit on an SM");
    int blocks = sms * per_sm;                        // all co-resident: required by the stripe wavefront
    const int want = (n_stripes + LONG_WARPS - 1) / LONG_WARPS;
    if (blocks > want) blocks = want;
    LongArgs a = args;
    void *params[] = {&a};
    AGX_CUDA(cudaLaunchCooperativeKernel((const void *)sw_longr_kernel<K, R, SHORT, DP4A, ENDS>, dim3(blocks), dim3(LONG_WARPS * 32),
                                         params, 0, st));
    count_launch();
    return AGX_OK;
}

template <int K, bool SHORT> int long_launch(const LongArgs &args, int n_stripes, cudaStream_t st)
{
    int dev = 0, sms = 0, per_sm = 0;
    AGX_CUDA(cudaGetDevice(&dev));
    AGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sw_long_kernel<K, SHORT>, LONG_WARPS * 32, 0));
    if (per_sm < 1) return fail(AGX_ECUDA, "sw_long: kernel does not fit on an SM");
    int blocks = sms * per_sm;                        // all co-resident: required by the stripe wavefront
    const int want = (n_stripes + LONG_WARPS - 1) / LONG_WARPS;
    if (blocks > want) blocks = want;
    LongArgs a = args;
    void *params[] = {&a};
    AGX_CUDA(cudaLaunchCooperativeKernel((const void *)sw_long_kernel<K, SHORT>, dim3(blocks), dim3(LONG_WARPS * 32),
                                         params, 0, st));
    count_launch();
    return AGX_OK;
}

int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return (e && *e && atoi(e) >= 0) ? atoi(e) : dflt;
}

constexpr int LONG_SLACK_DEFAULT = 0;
// row steps per hand-off block of sw_longr_kernel (the stripe-fill term of the run is stripes x (32 + B) steps)
int long_block_steps()
{
    // measured at 1 Mbp x 1 Mbp (profiles/r2v_sw_long_b.jsonl): 8 GPUs 148 / 134 / 133 ms for 32 / 16 / 8 steps (the
    // wavefront crosses 4464 stripes, each hop costs 31 + B steps), one GPU 350 / 344 / 355 ms
    const int b = env_int("AGX_LONG_B", 16);
    return b < 1 ? 1 : b > LR_BMAX ? LR_BMAX : b;
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

template <int I = 0> int long_dispatch_raw(int k, bool short_chain, const LongArgs &args, int n, cudaStream_t st)
{
    if constexpr (I < (int)(sizeof(LONG_KS_RAW) / sizeof(LONG_KS_RAW[0]))) {
        if (LONG_KS_RAW[I] == k)
            return short_chain ? long_launch<LONG_KS_RAW[I], true>(args, n, st) : long_launch<LONG_KS_RAW[I], false>(args, n, st);
        return long_dispatch_raw<I + 1>(k, short_chain, args, n, st);
    } else {
        return fail(AGX_EINVAL, "sw_long: stripe width not instantiated");
    }
}

template <int K, int R, bool DP4A> int long_launch_p(const LongArgs &args, int n_stripes, cudaStream_t st)
{
    int dev = 0, sms = 0, per_sm = 0;
    AGX_CUDA(cudaGetDevice(&dev));
    AGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sw_longp_kernel<K, R, DP4A>, LONG_WARPS * 32, 0));
    if (per_sm < 1) return fail(AGX_ECUDA, "sw_long: kernel does not fit on an SM");
    int blocks = sms * per_sm;
    const int want = (n_stripes + LONG_WARPS - 1) / LONG_WARPS;
    if (blocks > want) blocks = want;
    LongArgs a = args;
    void *params[] = {&a};
    AGX_CUDA(cudaLaunchCooperativeKernel((const void *)sw_longp_kernel<K, R, DP4A>, dim3(blocks), dim3(LONG_WARPS * 32),
                                         params, 0, st));
    count_launch();
    return AGX_OK;
}

// sw_longr_kernel instantiations: stripe widths x rows per step x chain form x substitution form
#ifdef AGX_LONG_SWEEP
constexpr int LONGR_KS[] = {6, 7, 8, 14, 27};
#define AGX_LONGR_VARIANTS(X) X(1, false, false) X(1, true, false) X(2, false, false) X(2, true, false) X(4, false, false) \
    X(4, true, false) X(1, false, true) X(1, true, true) X(2, false, true) X(2, true, true) X(4, false, true) X(4, true, true)
#define AGX_LONGP_VARIANTS(X) X(1, false) X(2, false) X(3, false) X(4, false) X(1, true) X(2, true) X(3, true) X(4, true) X(6, true)
#else
#define AGX_LONGP_VARIANTS(X) X(4, true)
constexpr int LONGR_KS[] = {2, 4, 6, 7, 8, 10, 12, 14, 16, 20, 24, 27, 32};
#define AGX_LONGR_VARIANTS(X) X(2, false, true) X(4, false, true)
#endif
constexpr int N_LONGR_KS = (int)(sizeof(LONGR_KS) / sizeof(LONGR_KS[0]));

template <int I = 0> int long_dispatch_r(int k, int rows, bool short_chain, bool dp4a, const LongArgs &args, int n, cudaStream_t st)
{
    if constexpr (I < N_LONGR_KS) {
        if (LONGR_KS[I] == k) {
#ifndef AGX_LONG_SWEEP
            if (args.best_key != nullptr) {      // END CELL wanted: the lean chain with the dp4a substitution
                if (rows == 2) return long_launch_r<LONGR_KS[I], 2, false, true, true>(args, n, st);
                return long_launch_r<LONGR_KS[I], 4, false, true, true>(args, n, st);
            }
#endif
#define AGX_X(RR, SS, DD) if (rows == RR && short_chain == SS && dp4a == DD) return long_launch_r<LONGR_KS[I], RR, SS, DD>(args, n, st);
            AGX_LONGR_VARIANTS(AGX_X)
#undef AGX_X
            return fail(AGX_EINVAL, "sw_long: rows-per-step / chain / dp4a combination not instantiated");
        }
        return long_dispatch_r<I + 1>(k, rows, short_chain, dp4a, args, n, st);
    } else {
        return fail(AGX_EINVAL, "sw_long: stripe width not instantiated");
    }
}
template <int I = 0> int long_dispatch_p(int k, int rows, bool dp4a, const LongArgs &args, int n, cudaStream_t st)
{
    if constexpr (I < N_LONGR_KS) {
        if (LONGR_KS[I] == k) {
#define AGX_X(RR, DD) if constexpr (RR <= LONGR_KS[I]) { if (rows == RR && dp4a == DD) return long_launch_p<LONGR_KS[I], RR, DD>(args, n, st); }
            AGX_LONGP_VARIANTS(AGX_X)
#undef AGX_X
            return fail(AGX_EINVAL, "sw_long: rows-per-step / dp4a combination not instantiated");
        }
        return long_dispatch_p<I + 1>(k, rows, dp4a, args, n, st);
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
    if (!args.lut) {
        if (args.best_key) return fail(AGX_ERANGE, "sw_long: end cells need sequences with at most 7 distinct bytes");
        return long_dispatch_raw<0>(k, short_chain, args, n, st);
    }
    // Measured on the per-GPU share of an 8-GPU run (125 kbp x 1 Mbp, one warp per scheduler) and on 1 Mbp x 1 Mbp
    // on one GPU (two warps per scheduler), profiles/r2a_long_sweep_*.jsonl: the lean chain with the dp4a
    // substitution wins everywhere; four rows per step when a scheduler holds a single warp (68 vs 92 ms), two
    // when it holds two (348 vs 388 ms).  The software-pipelined form is kept for experiments (AGX_LONG_PIPE=1).
    const int rows = env_int("AGX_LONG_R", n <= sms * 4 ? 4 : 2);
    const bool dp4a = env_int("AGX_LONG_DP4A", 1) != 0;
    if (getenv("AGX_LONG_CHAIN") == nullptr) short_chain = false;
    if (env_int("AGX_LONG_PIPE", 0) != 0 && rows <= k) return long_dispatch_p<0>(k, rows, dp4a, args, n, st);
    return long_dispatch_r<0>(k, rows, short_chain, dp4a, args, n, st);
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
    const int64_t lbp = lr_rows_padded(lb);
    // int32 words: entries, best, presence mask, code table, row score tables
    const int64_t need = 4 * lb + 4 + 8 + 64 + 2 * lbp + 4;
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
    LongArgs args = {};
    args.lut = coded ? d_lut : nullptr;
    args.one = 1;
    args.a = d_a; args.la = (int32_t)la; args.b = d_b; args.lb = (int32_t)lb;
    args.bsteps = long_block_steps();
    args.slack = env_int("AGX_LONG_SLACK", LONG_SLACK_DEFAULT);
    args.req_eighths = env_int("AGX_LONG_REQ", 4);
    int2 *d_rowtab = reinterpret_cast<int2 *>(ws.buf + (4 * lb + 4 + 8 + 64 + 1) / 2 * 2);
    args.rowtab = d_rowtab;
    if (coded) {
        long_rowtab_kernel<<<sms * 2, 256, 0, st>>>(d_b, (int32_t)lb, lbp, d_lut, sc, d_rowtab);
        count_launch();
    }
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
    ws.prof.destroy();
    ws = SwLongWorkspace();
}

// Host form over n_dev GPUs (device ordinals dev[], streams st[], workspaces ws[]): columns of `a`
// are split into n_dev contiguous ranges; `b` is replicated.  Returns the score in *score_out.
int sw_long_host_multi(int n_dev, const int *dev, cudaStream_t *st, SwLongWorkspace **ws, const uint8_t *a,
                       int64_t la, const uint8_t *b, int64_t lb, SwScoring sc, int32_t *score_out, int32_t *end_col_out,
                       int32_t *end_row_out, int col_is_sx)
{
    const bool want_ends = end_col_out != nullptr;
    if (want_ends) { *end_col_out = -1; *end_row_out = -1; }
    if (la <= 0 || lb <= 0) { *score_out = 0; return AGX_OK; }
    if (want_ends && (la > (int64_t)LONG_XMAX || lb > (int64_t)LONG_XMAX || (int64_t)std::min(la, lb) * sc.match > (int64_t)LONG_XMAX))
        return fail(AGX_ERANGE, "sw_long: end cells are kept for sequences up to 2^21 - 1 symbols");
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
        const int64_t lbp = lr_rows_padded(lb);
        const int64_t need = 4 * lb + 4 + 2 * lbp;         // one 16-byte entry per row + the best score + row score tables
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
        x.best_key = want_ends ? reinterpret_cast<unsigned long long *>(w.buf + 4 * lb + 2) : nullptr;
        x.col_base = (int32_t)c_lo[gidx];
        x.col_is_sx = col_is_sx;
        x.k32 = 32;
        x.next_bnd = nullptr;
        x.bsteps = long_block_steps();
        x.slack = env_int("AGX_LONG_SLACK", LONG_SLACK_DEFAULT);
        x.req_eighths = env_int("AGX_LONG_REQ", 4);
        int2 *d_rowtab = reinterpret_cast<int2 *>(w.buf + 4 * lb + 4);
        x.rowtab = d_rowtab;
        if (coded) {
            long_rowtab_kernel<<<sms * 2, 256, 0, st[gidx]>>>(x.b, (int32_t)lb, lbp, d_lut, sc, d_rowtab);
            count_launch();
        }
        x.stripe_base = stripe_base;
        stripe_base += (int32_t)n_stripes;
        x.sc = sc;
        AGX_CUDA(cudaMemsetAsync(x.bnd, 0xff, (size_t)lb * sizeof(int4), st[gidx]));   // tag -1: written by nobody
        AGX_CUDA(cudaMemsetAsync(x.best, 0, 4 * sizeof(int32_t), st[gidx]));      // the score and the end-cell key
    }
    for (int gidx = 0; gidx + 1 < n_dev; ++gidx) args[gidx].next_bnd = args[gidx + 1].bnd;
    for (int gidx = 0; gidx < n_dev; ++gidx) {
        AGX_CUDA(cudaSetDevice(dev[gidx]));
        AGX_CUDA(cudaStreamSynchronize(st[gidx]));
    }
    for (int gidx = 0; gidx < n_dev; ++gidx) {
        AGX_CUDA(cudaSetDevice(dev[gidx]));
        ws[gidx]->prof.begin(st[gidx]);
        int rc = long_dispatch(ks[gidx], args[gidx], st[gidx]);
        ws[gidx]->prof.end(st[gidx]);
        if (rc != AGX_OK) return rc;
    }
    int32_t best = 0;
    unsigned long long key = 0ull;
    for (int gidx = 0; gidx < n_dev; ++gidx) {
        AGX_CUDA(cudaSetDevice(dev[gidx]));
        int32_t v[4] = {0, 0, 0, 0};
        AGX_CUDA(cudaMemcpyAsync(v, args[gidx].best, sizeof v, cudaMemcpyDeviceToHost, st[gidx]));
        AGX_CUDA(cudaStreamSynchronize(st[gidx]));
        best = std::max(best, v[0]);
        key = std::max(key, (unsigned long long)(uint32_t)v[2] | ((unsigned long long)(uint32_t)v[3] << 32));
    }
    *score_out = best;
    if (want_ends && best > 0) {
        // key = H << 43 | (2^22 - 1 - (row + column)) << 21 | (2^21 - 1 - ix)
        const int64_t d = (int64_t)LONG_DMAX - (int64_t)((key >> 21) & LONG_DMAX);
        const int64_t x = (int64_t)LONG_XMAX - (int64_t)(key & LONG_XMAX);
        if ((int64_t)(key >> 43) != best) return fail(AGX_ECUDA, "sw_long: end-cell key and score disagree (internal error)");
        *end_col_out = (int32_t)(col_is_sx ? x : d - x);
        *end_row_out = (int32_t)(col_is_sx ? d - x : x);
    }
    return AGX_OK;
}

}  // namespace agx
