// sw_kernels.cu -- score-only affine-gap Smith-Waterman on B200 (sm_100a).
//
// Replaces the DP of smithWaterman/antidiagonalSmithWaterman.c:246-348 (reference paths are
// relative to the reference repository).  Recurrence restated:
//     P[i][j] = max(D[i-1][j] + go + ge, P[i-1][j] + ge)        vertical gap   ("F" below)
//     Q[i][j] = max(D[i][j-1] + go + ge, Q[i][j-1] + ge)        horizontal gap ("E" below)
//     D[i][j] = max(P, Q, D[i-1][j-1] + (x==y ? match : mismatch), 0)      ("H" below)
//     score   = max D
// with D = 0 on row/column 0 and P/Q = -inf there.  Because D >= 0, every P/Q value that is <= 0
// is interchangeable with -inf, so boundaries are initialised with goe = go + ge instead of INT_MIN.
//
// Three kernels:
//   sw_classify_kernel   device-side batch packer: strips the trailing '\n' symbol, picks the
//                        length class of every pair and appends it to that class's work list.
//   sw_duo_kernel<G,K>   INTER-TASK kernel.  A sub-warp of G lanes scores TWO pairs at once, one in
//                        each 16-bit half of a register (DPX s16x2: VIADDMNMX / VIMNMX3.RELU /
//                        VIADD.16x2).  Each lane keeps K columns of both pairs in registers and the
//                        rows stream through the sub-warp systolically (lane t is one row behind
//                        lane t-1; the boundary column moves with __shfl_up_sync).
//   sw_wave_kernel       INTRA-TASK kernel for everything else (long sequences, non-ACGT bytes,
//                        scores that could overflow s16): one warp per pair, s32 DPX, 32 lanes x 8
//                        columns per stripe, stripes chained through a boundary column in global
//                        memory.  A symbol-coded pass (A C G T N and newline: PRMT substitution score,
//                        additions on the FMA pipe) runs first; a pair with any other byte is redone by
//                        the raw-byte pass, so '\n', 'N', lower case ... behave exactly as in the
//                        reference (:332).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace agx {

namespace {

// ------------------------------------------------------------------------------------------
// class table of the duo kernel: columns capacity = G*K
// ------------------------------------------------------------------------------------------
struct DuoClass { int g, k; };
__host__ __device__ constexpr DuoClass duo_class(int c)
{
    // 256 and 512 columns: K = 16 on wider sub-warps instead of K = 32 (255 registers, two CTAs per SM) -- measured
    // 5598 -> 5783 and 5341 -> 5688 GCUPS (450-500 bp pairs: 4873 -> 5201); for 192 and 384 columns the narrower
    // sub-warp with K = 24 stays ahead of {16,12} / {32,12} (profiles/r2t_bench_*.json)
    // up to 64 columns: FOUR lanes per pair of pairs (3 steps of systolic skew instead of 7 on 24 .. 64 rows): 32 bp
    // 3208 -> 3813, 64 bp 4588 -> 5147 GCUPS (profiles/r2ao_bench_*.json)
    constexpr DuoClass t[SW_N_DUO_CLASSES] = {{4, 8},   {4, 16},  {8, 12},  {8, 16},  {8, 19}, {8, 24},
                                              {16, 16}, {16, 24}, {32, 16}, {32, 24}, {32, 32}};
    return t[c];
}
__host__ __device__ constexpr int duo_cap(int c) { return duo_class(c).g * duo_class(c).k; }
constexpr int DUO_MAX_CAP = duo_cap(SW_N_DUO_CLASSES - 1);  // 1024
constexpr int GENERIC = SW_N_DUO_CLASSES;
constexpr int LONGC = SW_N_DUO_CLASSES + 1;

// counters layout (int32): [0, NC) class counts / append cursors; NC: max raw length seen
constexpr int CNT_MAXLEN = SW_N_CLASSES;
constexpr int CNT_WORDS = SW_N_CLASSES + 4;

struct DuoConst {
    uint32_t goe2;    // (go+ge) in both halves
    uint32_t ext2;    // ge in both halves
    uint32_t xb4;     // byte (mismatch - goe) replicated 4x      : PRMT source for "no match"
    uint32_t mxor;    // (match - goe) ^ (mismatch - goe), one byte
    int32_t goe, match;
};

// prmt.b32 in its default mode: selector nibble bit 3 replicates the sign of the selected byte
// (__byte_perm masks that bit off, so the PTX is spelled out).
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}

__device__ __forceinline__ bool strip_newline(const uint8_t *p, int32_t &n)
{
    if (n > 0 && p[n - 1] == '\n') { --n; return true; }
    return false;
}

// ------------------------------------------------------------------------------------------
// classify / bin
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sw_classify_kernel(const uint8_t *__restrict__ seqs, const int64_t *__restrict__ off,
                   const int32_t *__restrict__ len, int64_t n_pairs, int32_t s16_max_short,
                   int64_t long_cells, int32_t match, int32_t *__restrict__ order, int32_t *__restrict__ counters,
                   int32_t *__restrict__ scores)
{
    __shared__ int32_t s_cnt[SW_N_CLASSES];
    __shared__ int32_t s_base[SW_N_CLASSES];
    __shared__ int32_t s_maxlen;
    if (threadIdx.x < SW_N_CLASSES) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_maxlen = 0;
    __syncthreads();

    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int cls = -1, rank = 0;
    if (p < n_pairs) {
        int32_t lx = len[2 * p], ly = len[2 * p + 1];
        const int32_t raw_max = lx > ly ? lx : ly;
        const bool nx = strip_newline(seqs + off[2 * p], lx);
        const bool ny = strip_newline(seqs + off[2 * p + 1], ly);
        const int32_t shorter = lx < ly ? lx : ly;
        if (shorter == 0) {
            // one side holds nothing but (possibly) its newline symbol: the only positive cell is
            // '\n' against '\n', present iff both lines end with one
            scores[p] = (nx && ny) ? match : 0;
        } else {
            cls = GENERIC;
            if ((int64_t)len[2 * p] * (int64_t)len[2 * p + 1] >= long_cells && shorter > DUO_MAX_CAP) cls = LONGC;
            else if (shorter <= s16_max_short) {
#pragma unroll
                for (int c = SW_N_DUO_CLASSES - 1; c >= 0; --c)
                    if (shorter <= duo_cap(c)) cls = c;
            }
            rank = atomicAdd(&s_cnt[cls], 1);
            atomicMax(&s_maxlen, raw_max);
        }
    }
    __syncthreads();
    if (threadIdx.x < SW_N_CLASSES && s_cnt[threadIdx.x] > 0)
        s_base[threadIdx.x] = atomicAdd(&counters[threadIdx.x], s_cnt[threadIdx.x]);
    if (threadIdx.x == 0 && s_maxlen > 0) atomicMax(&counters[CNT_MAXLEN], s_maxlen);
    __syncthreads();
    if (cls >= 0) order[(int64_t)cls * n_pairs + s_base[cls] + rank] = (int32_t)p;
}

// ------------------------------------------------------------------------------------------
// inter-task duo kernel (s16x2 DPX)
// ------------------------------------------------------------------------------------------
constexpr int DUO_THREADS = 128;
#ifndef AGX_DUO_CH
#define AGX_DUO_CH 32
#endif
// rows converted per refill (the storing mode keeps 32: its staging buffers need the shared memory) and row ring
// slots (>= rows per refill + G)
__host__ __device__ constexpr int duo_ch(int MODE) { return MODE == 2 ? 32 : AGX_DUO_CH; }

struct DuoSeq {
    const uint8_t *a;  // columns (shorter sequence)
    const uint8_t *b;  // rows (longer sequence)
    int32_t la, lb;
    bool both_nl;
};

__device__ __forceinline__ DuoSeq load_pair(const uint8_t *seqs, const int64_t *off,
                                            const int32_t *len, int32_t p)
{
    DuoSeq d;
    const uint8_t *x = seqs + off[2 * (int64_t)p];
    const uint8_t *y = seqs + off[2 * (int64_t)p + 1];
    int32_t lx = len[2 * (int64_t)p], ly = len[2 * (int64_t)p + 1];
    const bool nx = strip_newline(x, lx);
    const bool ny = strip_newline(y, ly);
    d.both_nl = nx && ny;
    if (lx <= ly) { d.a = x; d.la = lx; d.b = y; d.lb = ly; }
    else          { d.a = y; d.la = ly; d.b = x; d.lb = lx; }
    return d;
}

// Alignment modes keep the reference's orientation: ix (the columns here) walks line 2 only when line 1 is
// strictly longer, newline symbols counted (antidiagonalSmithWaterman.c:229-244) -- the end cell's tie rule is
// stated in that frame.  a_is_x: the columns are line 1.
__device__ __forceinline__ DuoSeq load_pair_oriented(const uint8_t *seqs, const int64_t *off,
                                                     const int32_t *len, int32_t p, bool &a_is_x)
{
    DuoSeq d;
    const uint8_t *x = seqs + off[2 * (int64_t)p];
    const uint8_t *y = seqs + off[2 * (int64_t)p + 1];
    int32_t lx = len[2 * (int64_t)p], ly = len[2 * (int64_t)p + 1];
    a_is_x = !(lx > ly);
    const bool nx = strip_newline(x, lx);
    const bool ny = strip_newline(y, ly);
    d.both_nl = nx && ny;
    if (a_is_x) { d.a = x; d.la = lx; d.b = y; d.lb = ly; }
    else        { d.a = y; d.la = ly; d.b = x; d.lb = lx; }
    return d;
}

// ASCII -> 2-bit code (A 0, C 1, T 2, G 3); ok is cleared for anything outside "ACGT".
__device__ __forceinline__ uint32_t base_code(uint32_t ch, bool &ok)
{
    const uint32_t code = (ch >> 1) & 3u;
    ok = ok && (((0x47544341u >> (8 * code)) & 0xffu) == ch);
    return code;
}

// Resident 128-thread CTAs per SM the register budget is cut for.  The state is 3 K registers (sel, Gp, F), so
// one budget for every class spills the wide ones (K = 32 at 5 CTAs / 96 registers: ~400 LDL/STL in the cell
// loop): 5 CTAs up to K = 19, 3 (168 registers) for K = 24, 2 (255) for K = 32.
#ifndef AGX_DUO_ALIGN_MINBLOCKS
#define AGX_DUO_ALIGN_MINBLOCKS 4
#endif
#ifndef AGX_DUO_MINBLOCKS
#define AGX_DUO_MINBLOCKS 5
#endif
// MODE 0: score only.  MODE 1: + END CELL.  MODE 2: + the low byte of every H, for the traceback walk.
// The alignment modes carry ~10 more live registers (keys, store pointer, packed bytes): one CTA fewer.
// MODE 2 matrix layout.  A warp's matrix is a sequence of BATCHES of TB steps; in a batch every lane owns a block of
// K2 x TB words, word (q, r) = columns 2q, 2q+1 (both pairs) of the lane's row of step r, stored at q * TB + r: the
// TB rows of a column pair sit side by side, so the traceback walk -- one row and one column back per step -- finds
// up to TB consecutive cells of its diagonal inside one 64-byte line.  Blocks are padded to an ODD number of words
// (the lanes' 32-bit shared-memory stores then fall into 32 different banks); the padding travels to HBM.
__host__ __device__ constexpr int duo_tb_steps(int K)
{
    const int k2 = (K + 1) / 2;                // a power of two that divides the 32-row refill; ~40 KB of staging per CTA
    return k2 <= 4 ? 8 : k2 <= 10 ? 4 : 2;
}
__host__ __device__ constexpr int duo_tb_block_words(int K) { return ((K + 1) / 2) * duo_tb_steps(K) + 1; }
__host__ __device__ constexpr int duo_tb_batch_bytes(int K) { return 32 * duo_tb_block_words(K) * 4; }
__host__ __device__ constexpr int duo_tb_smem_bytes(int K) { return (DUO_THREADS / 32) * 2 * duo_tb_batch_bytes(K); }
__host__ __device__ constexpr int duo_min_blocks_sw(int K, int MODE = 0)
{
    return K <= 19 ? (MODE ? AGX_DUO_ALIGN_MINBLOCKS : AGX_DUO_MINBLOCKS) : K <= 24 ? 3 : 2;
}
// what the alignment modes hand over besides the score
struct DuoAlignOut {
    int32_t *ends;        // [2 * n_pairs] end cell as (index in line 1, index in line 2), -1 -1 when the score is 0
    SwWalkRec *wk;        // [n_pairs]     MODE 2: where the traceback walk starts
    uint8_t *tb;          // MODE 2: H-byte matrices of this class, one per WARP, step-major:
                          //         [step s][lane][K2 words], lane t of a sub-warp holding row s - t at step s
    int64_t tb_class_off; //         where this class starts in the scratch (the walk addresses from the scratch base)
    int64_t tb_warp_bytes; //        batches * duo_tb_batch_bytes(K), batches = ceil(((most rows of the class) + G - 1) / TB)
    int32_t k32;          // 32, opaque to ptxas: key = H * 32 + tag as one IMAD on the FMA pipe
    int16_t cls;
};
template <int G, int K, int MODE>
__global__ void __launch_bounds__(DUO_THREADS, duo_min_blocks_sw(K, MODE))
sw_duo_kernel(const uint8_t *__restrict__ seqs, const int64_t *__restrict__ off,
              const int32_t *__restrict__ len, const int32_t *__restrict__ order_cls,
              int32_t n_in_class, DuoConst kc, int32_t *__restrict__ scores,
              int32_t *__restrict__ generic_list, int32_t *__restrict__ generic_cursor, DuoAlignOut ao)
{
    constexpr int CAP = G * K;
    constexpr int K2 = (K + 1) / 2;       // MODE 2: 32-bit words per thread row (2 columns x 2 pairs, one byte each)
    static_assert(K2 % 2 == 0, "thread rows are stored as 64-bit words");
    // MODE 2 stages TB steps of the whole warp in shared memory (TB * 32 * K2 words, contiguous in the step-major
    // global layout as well) and hands them to the copy engine as ONE bulk store; two buffers per warp
    constexpr int TB = duo_tb_steps(K);
    constexpr int LS = duo_tb_block_words(K);          // words of a lane's block in a batch
    extern __shared__ __align__(128) uint32_t tb_stage[];   // MODE 2: [warp][2 buffers][32 lanes][LS]
    constexpr int SUBS = DUO_THREADS / G;
    // row steps per loop trip: 8 measured +3 % over 2 at K = 19 (16 overflows the instruction cache: -17 %);
    // the wide classes keep 2
    // (MODE 2 adds ~35 % to a step: 8 of those no longer fit the instruction cache -- measured per 8 * 10^5 pairs
    // of 150 x 150: 8 -> 5.29 ms, 4 -> 4.61 ms, 2 -> 4.69 ms, 1 -> 4.75 ms; it unrolls its staged batch of TB steps)
    constexpr int STEP_UNROLL = (K <= 19) ? 8 : 2;
    constexpr int DUO_CH = duo_ch(MODE), DUO_RING = 2 * DUO_CH;
    __shared__ uint2 ring[SUBS][DUO_RING];

    const int lane = threadIdx.x & 31;
    const int t = threadIdx.x & (G - 1);          // lane inside the sub-warp
    const int sub = threadIdx.x / G;              // sub-warp inside the CTA
    const int duo = blockIdx.x * SUBS + sub;

    // ---- the two pairs of this sub-warp ------------------------------------------------------
    int32_t pid[2] = {-1, -1};
    DuoSeq sq[2];
    bool a_is_x[2] = {true, true};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int32_t slot = 2 * duo + h;
        sq[h].a = sq[h].b = nullptr; sq[h].la = sq[h].lb = 0; sq[h].both_nl = false;
        if (slot < n_in_class) {
            pid[h] = order_cls[slot];
            if constexpr (MODE == 0) sq[h] = load_pair(seqs, off, len, pid[h]);
            else sq[h] = load_pair_oriented(seqs, off, len, pid[h], a_is_x[h]);
        }
    }
    // ---- column selectors (columns right-aligned; padding columns select "sign of byte 0") ----
    // 'N' (any one extra symbol would do): as a ROW symbol it is the all-mismatch table, which is exact as long
    // as the column sequence holds no 'N'.  So a pair whose column sequence has an 'N' is swapped (the score is
    // symmetric) when the other sequence fits the columns and has none; only pairs with an 'N' on both sides,
    // or with any other byte outside ACGT, go to the byte-exact kernel.
    bool okh[2] = {true, true};      // per half: such a byte spoils only the pair that holds it
    uint32_t sel[K];
    const uint32_t submask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
#pragma unroll
    for (int j = 0; j < K; ++j) {
        uint32_t s = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int idx = t * K + j - (CAP - sq[h].la);
            uint32_t nib_lo = 8u + 4u * h, nib_hi = 8u + 4u * h;     // padding: 0x0000 / 0xffff
            if (idx >= 0) {
                const uint32_t code = base_code(sq[h].a[idx], okh[h]);
                nib_lo = code + 4u * h;
                nib_hi = nib_lo | 8u;                                 // sign-extend that byte
            }
            s |= (nib_lo | (nib_hi << 4)) << (8 * h);
        }
        sel[j] = s;
    }
    if (__any_sync(0xffffffffu, !(okh[0] && okh[1]))) {
        // cold path: some column sequence of this warp holds a byte outside ACGT.  Is it only 'N'?  Then swap
        // the pair when the other sequence fits the columns; the selectors are rebuilt (twice at most).
#pragma unroll 1
        for (int attempt = 0; attempt < 2; ++attempt) {
            bool saw_n[2] = {false, false};
            okh[0] = okh[1] = true;
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
                uint32_t s = 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int idx = t * K + j - (CAP - sq[h].la);
                    uint32_t nib_lo = 8u + 4u * h, nib_hi = 8u + 4u * h;
                    if (idx >= 0) {
                        const uint32_t ch = sq[h].a[idx];
                        bool valid = true;
                        const uint32_t code = base_code(ch, valid);
                        saw_n[h] = saw_n[h] || ch == 'N';
                        okh[h] = okh[h] && (valid || ch == 'N');
                        nib_lo = code + 4u * h;
                        nib_hi = nib_lo | 8u;
                    }
                    s |= (nib_lo | (nib_hi << 4)) << (8 * h);
                }
                // sel[] is indexed dynamically here (the loop is not unrolled to keep this path small)
#pragma unroll
                for (int jj = 0; jj < K; ++jj)
                    if (jj == j) sel[jj] = s;
            }
            bool redo = false;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (__ballot_sync(0xffffffffu, saw_n[h]) & submask) {   // uniform inside the sub-warp
                    // (the alignment modes never swap: their tie rule is tied to the orientation)
                    if (MODE == 0 && attempt == 0 && sq[h].lb <= CAP) {
                        const uint8_t *tp = sq[h].a; sq[h].a = sq[h].b; sq[h].b = tp;
                        const int32_t tl = sq[h].la; sq[h].la = sq[h].lb; sq[h].lb = tl;
                        redo = true;
                    } else {
                        okh[h] = false;                              // 'N' on both sides, or the other side does not fit
                    }
                }
            }
            if (!__any_sync(0xffffffffu, redo)) break;
        }
    }
    // rows are bottom-aligned to the longest b of the whole warp so the step loop is warp-uniform
    int32_t Lb = max(sq[0].lb, sq[1].lb);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) Lb = max(Lb, __shfl_xor_sync(0xffffffffu, Lb, m));
    const int S = Lb + G - 1;

    // ---- state ---------------------------------------------------------------------------------
    uint32_t Gp[K], F[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { Gp[j] = kc.goe2; F[j] = kc.goe2; }
    uint32_t rmax = kc.goe2, g_out = kc.goe2, e_out = kc.goe2, g_in_prev = kc.goe2;
    // Alignment modes.  Every cell gets a 15-bit key H * 32 + (31 - j) per half (one IMAD for both halves; H <= 1023),
    // the row maximum of the keys names the row's best H and its FIRST column, and once per row that is widened
    // to a 32-bit key that orders cells as the reference visits them:
    //     H << 22 | (4095 - d) << 10 | (1023 - c),   c = kernel column, d = kernel row + c + 31
    // (larger H first, then the earlier anti-diagonal, then the smaller ix).  best32 = max of those keys.
    uint32_t best32[2] = {0u, 0u};
    uint32_t kbase = ((uint32_t)(4033 - t * K + t) << 10) + (uint32_t)(992 - t * K);   // row -t: the first step of lane t
    const int wib = threadIdx.x >> 5;
    uint8_t *tb_warp = nullptr;               // this warp's matrix
    int tb_cur = 0;                           // staging buffer in use
    int64_t tb_done = 0;                      // bytes handed to the copy engine so far
    if constexpr (MODE == 2) tb_warp = ao.tb + ((int64_t)blockIdx.x * (DUO_THREADS / 32) + wib) * ao.tb_warp_bytes;
    // hand the staged steps to the copy engine: the writes of every lane become visible to the async proxy, one
    // lane issues the bulk store and makes sure the store that read the OTHER buffer (a batch ago) is done reading
    auto tb_flush = [&](const int n_steps) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)duo_tb_batch_bytes(K);     // always a whole batch (unused steps: stale words)
            const uint32_t src = (uint32_t)__cvta_generic_to_shared(tb_stage + (wib * 2 + tb_cur) * 32 * LS);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(tb_warp + tb_done), "r"(src), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncwarp();
        (void)n_steps;
        tb_done += duo_tb_batch_bytes(K);
        tb_cur ^= 1;
    };

    for (int i = t; i < DUO_RING; i += G) ring[sub][i] = make_uint2(kc.xb4, kc.xb4);

    for (int s0 = 0; s0 < S; s0 += DUO_CH) {
        __syncwarp();
        // refill: rows [s0, s0 + DUO_CH) -> PRMT source words, 4 bytes per pair
#pragma unroll
        for (int i = 0; i < DUO_CH / G; ++i) {
            const int r = s0 + t + G * i;
            uint32_t w[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int idx = r - (Lb - sq[h].lb);
                w[h] = kc.xb4;
                if (idx >= 0 && r < Lb) {
                    const uint32_t ch = sq[h].b[idx];
                    if (ch != 'N') {                               // 'N': matches no column symbol (none is 'N')
                        const uint32_t code = base_code(ch, okh[h]);
                        w[h] = kc.xb4 ^ (kc.mxor << (8 * code));
                    }
                }
            }
            ring[sub][r & (DUO_RING - 1)] = make_uint2(w[0], w[1]);
        }
        __syncwarp();

        const int send = min(DUO_CH, S - s0);
        // one row step of every lane; slot = position of the step in the staged batch (MODE 2)
        auto step = [&](const int s, const int slot) {
            const uint2 R = ring[sub][(s - t) & (DUO_RING - 1)];
            uint32_t g_in = __shfl_up_sync(0xffffffffu, g_out, 1, G);
            uint32_t e = __shfl_up_sync(0xffffffffu, e_out, 1, G);
            if (t == 0) { g_in = kc.goe2; e = kc.goe2; }
            uint32_t gdiag = g_in_prev;
            g_in_prev = g_in;
            uint32_t gleft = g_in;
            uint32_t rk = 0u, kprev = 0u;
            uint32_t pk[MODE == 2 ? K2 : 1];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const uint32_t tt = prmt(R.x, R.y, sel[j]);                  // (subst - goe) per half
                const uint32_t d = __vadd2(gdiag, tt);                    // H[i-1][j-1] + subst
                e = __viaddmax_s16x2(e, kc.ext2, gleft);                  // E[i][j]
                F[j] = __viaddmax_s16x2(F[j], kc.ext2, Gp[j]);            // F[i][j]
                const uint32_t hcell = __vimax3_s16x2_relu(e, F[j], d);   // H[i][j]
                gdiag = Gp[j];
                gleft = __vadd2(hcell, kc.goe2);                          // H[i][j] + goe
                Gp[j] = gleft;
                if constexpr (MODE == 0) {
                    if (j & 1) rmax = __vimax3_s16x2(rmax, Gp[j - 1], gleft);
                    else if (j == K - 1) rmax = __vmaxs2(rmax, gleft);
                } else {
                    uint32_t key;
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(key) : "r"(hcell), "r"((uint32_t)ao.k32), "r"((uint32_t)(31 - j) * 0x00010001u));
                    if (j & 1) rk = __vimax3_s16x2(rk, kprev, key);
                    else if (j == K - 1) rk = __vmaxs2(rk, key);
                    kprev = key;
                }
                if constexpr (MODE == 2) {
                    // bytes [lo pair col j-1, lo pair col j, hi pair col j-1, hi pair col j] of H + goe: the walk uses
                    // differences of neighbouring bytes only, so the constant does not matter
                    if (j & 1) pk[j >> 1] = prmt(Gp[j - 1], gleft, 0x6240u);
                    else if (j == K - 1) pk[j >> 1] = prmt(gleft, 0u, 0x6240u);
                }
            }
            g_out = gleft;
            e_out = e;
            if constexpr (MODE != 0) {
                const uint32_t m = rk & 0xffe0ffe0u;             // H << 5 per half
                const uint32_t jj = rk & 0x001f001fu;            // 31 - (first column with that H) per half
                uint32_t k0, k1;
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(k0) : "r"(jj & 0xffffu), "r"(1025u), "r"(kbase));
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(k1) : "r"(jj >> 16), "r"(1025u), "r"(kbase));
                best32[0] = max(best32[0], k0 + (m << 17));
                best32[1] = max(best32[1], k1 + ((m >> 16) << 17));
                kbase -= 1024u;
            }
            if constexpr (MODE == 2) {
                uint32_t *dst = tb_stage + ((wib * 2 + tb_cur) * 32 + lane) * LS + slot;
#pragma unroll
                for (int q = 0; q < K2; ++q) dst[q * TB] = pk[q];
            }
        };
        if constexpr (MODE == 2) {
            // batches of TB steps, each handed to the copy engine as one bulk store
            for (int u0 = 0; u0 < send; u0 += TB) {
                const int ue = min(TB, send - u0);
                if (ue == TB) {
#pragma unroll
                    for (int k = 0; k < TB; ++k) step(s0 + u0 + k, k);
                } else {
#pragma unroll 1
                    for (int k = 0; k < ue; ++k) step(s0 + u0 + k, k);
                }
                tb_flush(ue);
            }
        } else {
#pragma unroll STEP_UNROLL
            for (int u = 0; u < send; ++u) step(s0 + u, 0);
        }
    }

    // ---- results ---------------------------------------------------------------------------------
    if constexpr (MODE == 2) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
#pragma unroll
    for (int m = G / 2; m >= 1; m >>= 1) rmax = __vmaxs2(rmax, __shfl_xor_sync(0xffffffffu, rmax, m, G));
    if constexpr (MODE != 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int m = G / 2; m >= 1; m >>= 1) best32[h] = max(best32[h], __shfl_xor_sync(0xffffffffu, best32[h], m, G));
    }
    const uint32_t corner2 = __shfl_sync(0xffffffffu, Gp[K - 1], G - 1, G);
    const uint32_t okbits0 = __ballot_sync(0xffffffffu, okh[0]), okbits1 = __ballot_sync(0xffffffffu, okh[1]);
    const bool all_ok[2] = {(okbits0 & submask) == submask, (okbits1 & submask) == submask};
    if (t == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (pid[h] < 0) continue;
            if (!all_ok[h]) {
                // a byte outside ACGT in this pair: let the byte-exact kernel redo it (the other half of
                // the registers never mixes with this one, so the duo's other pair keeps its score)
                generic_list[atomicAdd(generic_cursor, 1)] = pid[h];
                continue;
            }
            const int32_t corner = (int32_t)(int16_t)(corner2 >> (16 * h)) - kc.goe;
            if constexpr (MODE == 0) {
                const int32_t best = (int32_t)(int16_t)(rmax >> (16 * h)) - kc.goe;
                // '\n' is the last symbol of both lines: it can only match at the corner cell
                scores[pid[h]] = sq[h].both_nl ? max(best, corner + kc.match) : best;
            } else {
                const uint32_t key = best32[h];
                int32_t best = (int32_t)(key >> 22);
                const int32_t c_end = 1023 - (int32_t)(key & 1023u);
                const int32_t r_end = 4095 - (int32_t)((key >> 10) & 4095u) - c_end - 31;
                // kernel row / column of symbol 0 of the row / column sequence
                const int32_t row_off = Lb - sq[h].lb, col_off = CAP - sq[h].la;
                int32_t ea = c_end - col_off, eb = r_end - row_off;     // in the column / row sequence
                // the newline corner is the LAST cell the reference visits: it takes the maximum only when larger
                const bool nl_end = sq[h].both_nl && corner + kc.match > best;
                if (nl_end) { best = corner + kc.match; ea = sq[h].la; eb = sq[h].lb; }
                if (best == 0) ea = eb = -1;
                scores[pid[h]] = best;
                ao.ends[2 * (int64_t)pid[h]] = a_is_x[h] ? ea : eb;
                ao.ends[2 * (int64_t)pid[h] + 1] = a_is_x[h] ? eb : ea;
                if constexpr (MODE == 2) {
                    SwWalkRec w;
                    // the warp's matrix + this sub-warp's lanes inside a step
                    w.tb_off = ao.tb_class_off + ((int64_t)blockIdx.x * (DUO_THREADS / 32) + wib) * ao.tb_warp_bytes +
                               (int64_t)(lane & ~(G - 1)) * LS * 4;
                    w.r_end = nl_end ? Lb - 1 : r_end;
                    w.c_end = nl_end ? CAP - 1 : c_end;
                    w.row_off = row_off;
                    w.col_off = col_off;
                    w.rstride = 0;
                    w.cls = ao.cls;
                    w.half = (uint8_t)h;
                    w.flags = (uint8_t)((a_is_x[h] ? SW_WK_A_IS_X : 0) | (nl_end ? SW_WK_NL_END : 0) | (best == 0 ? SW_WK_NONE : 0));
                    ao.wk[pid[h]] = w;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// intra-task wavefront kernel (s32 DPX), one warp per pair, raw bytes
// ------------------------------------------------------------------------------------------
constexpr int WAVE_K = 8;                 // columns per lane
constexpr int WAVE_W = 32 * WAVE_K;       // stripe width
constexpr int WAVE_WARPS = 4;             // warps per CTA
constexpr int WAVE_RING = 64;

// prmt.b32 with a sign-extending selector, as a signed value
__device__ __forceinline__ int32_t prmt_s(uint32_t a, uint32_t b, uint32_t sel) { return (int32_t)prmt(a, b, sel); }
// a + b as IMAD (FMA pipe): the ALU pipe is what these kernels run out of; `one` is 1, opaque to ptxas
__device__ __forceinline__ int32_t add_fma(int32_t a, int32_t b, int32_t one)
{
    int32_t d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
}
// the CODED pass knows A C G T N '\n' (codes 0..5); code 7 is the "matches nothing" padding
__device__ __forceinline__ uint32_t wave_code(uint32_t ch, bool &ok)
{
    const uint32_t c2 = (ch >> 1) & 3u;
    if (((0x47544341u >> (8 * c2)) & 0xffu) == ch) return c2;
    if (ch == 'N') return 4u;
    if (ch == '\n') return 5u;
    ok = false;
    return 6u;
}

// One pair on one warp.  CODED: substitution score = one sign-extending PRMT from an 8-byte per-row table,
// additions on the FMA pipe (4.5 instead of 7.5 ALU-pipe instructions per cell, as in sw_long.cu); it gives
// up (returns false) when it meets a byte outside its alphabet and the raw-byte pass redoes the pair.
// MODE as in sw_duo_kernel.  The alignment modes order cells with a 64-bit key,
//     H << 43 | (2^29 - 1 - (row + column)) << 14 | (16383 - column)
// (H < 2^21, row + column < 2^29, column < 2^14: checked on the host); MODE 2 stores the low byte of every H at
// tb[((stripe * lb + row) * 32 + lane) * 8 + j].
constexpr uint32_t WAVE_DMAX = (1u << 29) - 1u;
template <bool CODED, int MODE>
__device__ bool wave_pair(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb, SwScoring sc, int32_t one,
                          int32_t *bnd, int32_t (*r_byte)[WAVE_RING], int32_t (*r_hi)[WAVE_RING],
                          int32_t (*r_g)[WAVE_RING], int32_t (*r_e)[WAVE_RING], int2 (*stage)[32], int32_t &best_out,
                          unsigned long long &key_out, uint8_t *tb, int32_t k32)
{
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int32_t goe = sc.gap_open + sc.gap_extend;
    const int32_t ext = sc.gap_extend;
    const int32_t sub_match = sc.match - goe, sub_mis = sc.mismatch - goe;
    const uint32_t xb4 = (uint32_t)(uint8_t)(int8_t)sub_mis * 0x01010101u;
    const uint32_t mxor = (uint32_t)(uint8_t)(int8_t)sub_mis ^ (uint32_t)(uint8_t)(int8_t)sub_match;
    bool ok = true;
    int32_t bestg = goe;
    unsigned long long best64 = 0ull;
    const int n_stripes = (la + WAVE_W - 1) / WAVE_W;
    for (int st = 0; st < n_stripes; ++st) {
        const int c0 = st * WAVE_W + lane * WAVE_K;
        int32_t acol[WAVE_K], Gp[WAVE_K], F[WAVE_K];
#pragma unroll
        for (int j = 0; j < WAVE_K; ++j) {
            if constexpr (CODED) {
                const uint32_t c = (c0 + j < la) ? wave_code(a[c0 + j], ok) : 7u;
                acol[j] = (int32_t)(c | ((8u | c) * 0x1110u));
            } else {
                acol[j] = (c0 + j < la) ? (int32_t)a[c0 + j] : 0x100;   // 0x100 never equals a byte
            }
            Gp[j] = goe;
            F[j] = goe;
        }
        int32_t g_out = goe, e_out = goe, g_in_prev = goe;
        const bool first = (st == 0), last = (st == n_stripes - 1);
        const int S = lb + 31;
        // boundary of the previous stripe for the first 32 rows; later blocks are requested a block ahead
        int2 nx = make_int2(goe, goe);
        if (!first && lane < lb) nx = __ldcg(reinterpret_cast<const int2 *>(bnd) + lane);
        for (int s0 = 0; s0 < S; s0 += 32) {
            __syncwarp();
            {
                const int r = s0 + lane;
                int32_t bb = CODED ? (int32_t)xb4 : 0x200;
                uint32_t hi = xb4;
                if (r < lb) {
                    if constexpr (CODED) {
                        const uint32_t c = wave_code(b[r], ok);
                        uint32_t lo = xb4;
                        if (c < 4) lo ^= mxor << (8 * c); else hi ^= mxor << (8 * (c - 4));
                        bb = (int32_t)lo;
                    } else {
                        bb = b[r];
                    }
                }
                r_byte[wib][r & (WAVE_RING - 1)] = bb;
                if constexpr (CODED) r_hi[wib][r & (WAVE_RING - 1)] = (int32_t)hi;
                r_g[wib][r & (WAVE_RING - 1)] = nx.x;
                r_e[wib][r & (WAVE_RING - 1)] = nx.y;
                // the rows this stripe overwrites in this block are s0-31 .. s0: rows s0+32.. are still the
                // previous stripe's, so the next block's boundary can be requested now
                const int rn = s0 + 32 + lane;
                nx = make_int2(goe, goe);
                if (!first && rn < lb) nx = __ldcg(reinterpret_cast<const int2 *>(bnd) + rn);
            }
            __syncwarp();
            const int send = min(32, S - s0);
            // row steps per loop trip: measured on 3 kbp pairs 1 -> 1766, 2 -> 1786, 4 -> 1874 GCUPS
#pragma unroll 4
            for (int u = 0; u < send; ++u) {
                const int s = s0 + u;
                const int slot = (s - lane) & (WAVE_RING - 1);
                const int32_t rb = (s - lane >= 0) ? r_byte[wib][slot] : (CODED ? (int32_t)xb4 : 0x200);
                uint32_t rhi = xb4;
                if constexpr (CODED) { if (s - lane >= 0) rhi = (uint32_t)r_hi[wib][slot]; }
                int32_t g_in = __shfl_up_sync(0xffffffffu, g_out, 1);
                int32_t e = __shfl_up_sync(0xffffffffu, e_out, 1);
                if (lane == 0) { g_in = r_g[wib][slot]; e = r_e[wib][slot]; }
                int32_t gdiag = g_in_prev;
                g_in_prev = g_in;
                int32_t gleft = g_in;
                int32_t rk = 0, kprev = 0;
                uint32_t hb[MODE == 2 ? WAVE_K : 1];
#pragma unroll
                for (int j = 0; j < WAVE_K; ++j) {
                    int32_t d;
                    if constexpr (CODED) d = add_fma(gdiag, prmt_s((uint32_t)rb, rhi, (uint32_t)acol[j]), one);
                    else d = gdiag + ((acol[j] == rb) ? sub_match : sub_mis);
                    e = __viaddmax_s32(e, ext, gleft);
                    F[j] = __viaddmax_s32(F[j], ext, Gp[j]);
                    const int32_t hcell = __vimax3_s32_relu(e, F[j], d);
                    gdiag = Gp[j];
                    if constexpr (CODED) gleft = add_fma(hcell, goe, one); else gleft = hcell + goe;
                    Gp[j] = gleft;
                    if constexpr (MODE == 0) {
                        if (j & 1) bestg = __vimax3_s32(bestg, Gp[j - 1], gleft);
                        else if (j == WAVE_K - 1) bestg = max(bestg, gleft);
                    } else {
                        int32_t key;                                   // H * 32 + (31 - j): the row maximum names H and its first column
                        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(key) : "r"(hcell), "r"(k32), "r"(31 - j));
                        if (j & 1) rk = __vimax3_s32(rk, kprev, key);
                        else if (j == WAVE_K - 1) rk = max(rk, key);
                        kprev = key;
                    }
                    if constexpr (MODE == 2) hb[j] = (uint32_t)hcell;
                }
                g_out = gleft;
                e_out = e;
                if constexpr (MODE != 0) {
                    const int r = s - lane;
                    const uint32_t col = (uint32_t)(c0 + 31 - (rk & 31));
                    const uint32_t dd = WAVE_DMAX - ((uint32_t)r + col);        // junk rows (r < 0) have H = 0
                    const unsigned long long k64 = ((unsigned long long)(uint32_t)(rk >> 5) << 43) |
                                                   ((unsigned long long)(dd & WAVE_DMAX) << 14) | (16383u - col);
                    best64 = max(best64, k64);
                    if constexpr (MODE == 2) {
                        if ((unsigned)r < (unsigned)lb) {
                            const uint32_t w0 = prmt(prmt(hb[0], hb[1], 0x0040u), prmt(hb[2], hb[3], 0x0040u), 0x5410u);
                            const uint32_t w1 = prmt(prmt(hb[4], hb[5], 0x0040u), prmt(hb[6], hb[7], 0x0040u), 0x5410u);
                            reinterpret_cast<uint2 *>(tb)[((int64_t)st * lb + r) * 32 + lane] = make_uint2(w0, w1);
                        }
                    }
                }
                if (lane == 31) stage[wib][u] = make_int2(g_out, e_out);     // row s - 31 of the last column
            }
            // one coalesced store per block instead of a global store per row step
            if (!last) {
                __syncwarp();
                const int r = s0 - 31 + lane;
                if (lane < send && r >= 0 && r < lb) reinterpret_cast<int2 *>(bnd)[r] = stage[wib][lane];
            }
        }
        __syncwarp();
        __threadfence_block();
        if constexpr (CODED) {
            if (!__all_sync(0xffffffffu, ok)) return false;     // a byte outside the coded alphabet: redo raw
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) bestg = max(bestg, __shfl_xor_sync(0xffffffffu, bestg, m));
    best_out = max(bestg - goe, 0);
    if constexpr (MODE != 0) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) best64 = max(best64, __shfl_xor_sync(0xffffffffu, best64, m));
        key_out = best64;
        best_out = (int32_t)(best64 >> 43);
    }
    return true;
}

struct WaveAlignOut {
    int32_t *ends;             // [2 * n_pairs]
    SwWalkRec *wk;             // [n_pairs]   MODE 2
    uint8_t *tb;               // MODE 2: H-byte matrices of the listed pairs
    const int64_t *tb_off;     //         [list position] byte offset of that pair's matrix
    int32_t k32;
};
template <int MODE>
__global__ void __launch_bounds__(WAVE_WARPS * 32)
sw_wave_kernel(const uint8_t *__restrict__ seqs, const int64_t *__restrict__ off,
               const int32_t *__restrict__ len, const int32_t *__restrict__ list,
               const int32_t *__restrict__ list_count, SwScoring sc, int32_t one, int32_t coded_ok,
               int32_t *__restrict__ scores, int32_t *__restrict__ scratch, int64_t scratch_stride, WaveAlignOut ao)
{
    __shared__ int32_t r_byte[WAVE_WARPS][WAVE_RING];
    __shared__ int32_t r_hi[WAVE_WARPS][WAVE_RING];
    __shared__ int32_t r_g[WAVE_WARPS][WAVE_RING];
    __shared__ int32_t r_e[WAVE_WARPS][WAVE_RING];
    __shared__ int2 stage[WAVE_WARPS][32];

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * WAVE_WARPS + wib;
    const int64_t n_warps = (int64_t)gridDim.x * WAVE_WARPS;
    int32_t *bnd = scratch + warp * scratch_stride;   // [2 * rows] : (G, E) of the stripe's last column
    const int32_t count = *list_count;

    for (int64_t it = warp; it < count; it += n_warps) {
        const int32_t p = list[it];
        const uint8_t *x = seqs + off[2 * (int64_t)p];
        const uint8_t *y = seqs + off[2 * (int64_t)p + 1];
        int32_t lx = len[2 * (int64_t)p], ly = len[2 * (int64_t)p + 1];
        // columns = shorter sequence (fewer stripes), rows = longer
        const uint8_t *a = x, *b = y;
        int32_t la = lx, lb = ly;
        if (lx > ly) { a = y; la = ly; b = x; lb = lx; }

        int32_t best = 0;
        unsigned long long key = 0ull;
        uint8_t *tb = nullptr;
        if constexpr (MODE == 2) tb = ao.tb + ao.tb_off[it];
        bool done = false;
        if (coded_ok) done = wave_pair<true, MODE>(a, la, b, lb, sc, one, bnd, r_byte, r_hi, r_g, r_e, stage, best, key, tb, ao.k32);
        if (!done) wave_pair<false, MODE>(a, la, b, lb, sc, one, bnd, r_byte, r_hi, r_g, r_e, stage, best, key, tb, ao.k32);
        if (lane == 0) {
            scores[p] = best;
            if constexpr (MODE != 0) {
                // the wavefront kernel sees raw bytes (newline symbols included) in the reference's orientation
                const int32_t col = 16383 - (int32_t)(key & 16383u);
                const int32_t row = (int32_t)(WAVE_DMAX - (uint32_t)((key >> 14) & WAVE_DMAX)) - col;
                const bool a_is_x = !(lx > ly);
                const int32_t ea = best > 0 ? col : -1, eb = best > 0 ? row : -1;
                ao.ends[2 * (int64_t)p] = a_is_x ? ea : eb;
                ao.ends[2 * (int64_t)p + 1] = a_is_x ? eb : ea;
                if constexpr (MODE == 2) {
                    SwWalkRec w;
                    w.tb_off = ao.tb_off[it];
                    w.r_end = row;
                    w.c_end = col;
                    w.row_off = 0;
                    w.col_off = 0;
                    w.rstride = lb;
                    w.cls = -1;
                    w.half = 0;
                    w.flags = (uint8_t)((a_is_x ? SW_WK_A_IS_X : 0) | SW_WK_RAW | (best == 0 ? SW_WK_NONE : 0));
                    ao.wk[p] = w;
                }
            }
        }
    }
}

template <int C>
int launch_duo_class(const uint8_t *d_seqs, const int64_t *d_off, const int32_t *d_len,
                     const int32_t *order, int64_t n_pairs, int32_t count, const DuoConst &kc,
                     int32_t *d_scores, int32_t *counters, cudaStream_t st)
{
    constexpr int G = duo_class(C).g, K = duo_class(C).k;
    constexpr int SUBS = DUO_THREADS / G;
    const int duos = (count + 1) / 2;
    const int blocks = (duos + SUBS - 1) / SUBS;
    if (blocks == 0) return AGX_OK;
    sw_duo_kernel<G, K, 0><<<blocks, DUO_THREADS, 0, st>>>(
        d_seqs, d_off, d_len, order + (int64_t)C * n_pairs, count, kc, d_scores,
        const_cast<int32_t *>(order) + (int64_t)GENERIC * n_pairs, counters + GENERIC, DuoAlignOut{});
    count_launch();
    AGX_CUDA(cudaGetLastError());
    return AGX_OK;
}

#include "sw_align.cuh"

template <int C>
int launch_all_duo(const uint8_t *d_seqs, const int64_t *d_off, const int32_t *d_len,
                   const int32_t *order, int64_t n_pairs, const int32_t *counts, const DuoConst &kc,
                   int32_t *d_scores, int32_t *counters, cudaStream_t st)
{
    if constexpr (C < SW_N_DUO_CLASSES) {
        int rc = launch_duo_class<C>(d_seqs, d_off, d_len, order, n_pairs, counts[C], kc, d_scores,
                                     counters, st);
        if (rc != AGX_OK) return rc;
        return launch_all_duo<C + 1>(d_seqs, d_off, d_len, order, n_pairs, counts, kc, d_scores,
                                     counters, st);
    } else {
        return AGX_OK;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side of the SW path
// ------------------------------------------------------------------------------------------
int64_t sw_long_cells()
{
    static const int64_t v = [] {
        const char *e = getenv("AGX_SW_LONG_CELLS");
        const long long x = e ? atoll(e) : 0;
        return x > 0 ? (int64_t)x : SW_LONG_CELLS_DEFAULT;
    }();
    return v;
}

int sw_workspace_reserve(SwWorkspace &ws, int64_t n_pairs)
{
    if (!ws.counters) {
        AGX_CUDA(cudaMalloc(&ws.counters, CNT_WORDS * sizeof(int32_t)));
        AGX_CUDA(cudaMallocHost(&ws.h_counters, CNT_WORDS * sizeof(int32_t)));
    }
    if (n_pairs > ws.cap_pairs) {
        if (ws.order) cudaFree(ws.order);
        ws.order = nullptr;
        ws.cap_pairs = 0;
        AGX_CUDA(cudaMalloc(&ws.order, (size_t)n_pairs * SW_N_CLASSES * sizeof(int32_t)));
        ws.cap_pairs = n_pairs;
    }
    return AGX_OK;
}

void sw_workspace_free(SwWorkspace &ws)
{
    if (ws.order) cudaFree(ws.order);
    if (ws.counters) cudaFree(ws.counters);
    if (ws.wave_scratch) cudaFree(ws.wave_scratch);
    if (ws.h_counters) cudaFreeHost(ws.h_counters);
    sw_long_workspace_free(ws.lng);
    ws.prof_duo.destroy(); ws.prof_wave.destroy(); ws.prof_classify.destroy(); ws.prof_long.destroy();
    ws = SwWorkspace();
}

int sw_run_device(SwWorkspace &ws, const uint8_t *d_seqs, const int64_t *d_off, const int32_t *d_len,
                  int64_t n_pairs, SwScoring sc, int32_t *d_scores, cudaStream_t st, cudaStream_t prep_st)
{
    if (n_pairs == 0) return AGX_OK;
    cudaStream_t cst = prep_st ? prep_st : st;     // stream of the length-class pass
    if (n_pairs > (int64_t)1 << 30) return fail(AGX_ERANGE, "sw: more than 2^30 pairs in one call");
    if (!(sc.match > 0 && sc.mismatch < 0 && sc.gap_open <= 0 && sc.gap_extend < 0))
        return fail(AGX_ERANGE, "sw: scoring must satisfy match > 0 > mismatch, gap_open <= 0, gap_extend < 0");
    int rc = sw_workspace_reserve(ws, n_pairs);
    if (rc != AGX_OK) return rc;

    const int32_t goe = sc.gap_open + sc.gap_extend;
    // s16x2 path: substitution bytes must fit a signed byte and the best score a signed 16-bit half
    const bool s16_ok = (sc.match - goe) <= 127 && (sc.mismatch - goe) >= -128 && goe >= -1024 &&
                        sc.match <= 30;
    const int32_t s16_max_short = s16_ok ? min(DUO_MAX_CAP, 32000 / sc.match) : 0;

    AGX_CUDA(cudaMemsetAsync(ws.counters, 0, CNT_WORDS * sizeof(int32_t), cst));
    const int cblocks = (int)((n_pairs + 255) / 256);
    ws.prof_classify.begin(cst);
    sw_classify_kernel<<<cblocks, 256, 0, cst>>>(d_seqs, d_off, d_len, n_pairs, s16_max_short,
                                                 sw_long_cells(), sc.match, ws.order, ws.counters, d_scores);
    count_launch();
    ws.prof_classify.end(cst);
    AGX_CUDA(cudaGetLastError());
    AGX_CUDA(cudaMemcpyAsync(ws.h_counters, ws.counters, CNT_WORDS * sizeof(int32_t),
                             cudaMemcpyDeviceToHost, cst));
    AGX_CUDA(cudaStreamSynchronize(cst));

    int32_t counts[SW_N_CLASSES];
    int64_t n_duo = 0;
    for (int c = 0; c < SW_N_CLASSES; ++c) counts[c] = ws.h_counters[c];
    for (int c = 0; c < SW_N_DUO_CLASSES; ++c) n_duo += counts[c];
    const int32_t max_len = ws.h_counters[CNT_MAXLEN];

    DuoConst kc;
    kc.goe = goe;
    kc.match = sc.match;
    kc.goe2 = ((uint32_t)(uint16_t)(int16_t)goe) * 0x00010001u;
    kc.ext2 = ((uint32_t)(uint16_t)(int16_t)sc.gap_extend) * 0x00010001u;
    const uint32_t xb = (uint32_t)(uint8_t)(int8_t)(sc.mismatch - goe);
    const uint32_t mb = (uint32_t)(uint8_t)(int8_t)(sc.match - goe);
    kc.xb4 = xb * 0x01010101u;
    kc.mxor = xb ^ mb;

    ws.prof_duo.begin(st);
    rc = launch_all_duo<0>(d_seqs, d_off, d_len, ws.order, n_pairs, counts, kc, d_scores,
                           ws.counters, st);
    ws.prof_duo.end(st);
    if (rc != AGX_OK) return rc;

    // whole-GPU alignments, one after the other (each one fills the device by itself)
    if (counts[LONGC] > 0) {
        std::vector<int32_t> ids(counts[LONGC]);
        AGX_CUDA(cudaMemcpyAsync(ids.data(), ws.order + (int64_t)LONGC * n_pairs, ids.size() * sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, st));
        AGX_CUDA(cudaStreamSynchronize(st));
        ws.prof_long.begin(st);
        for (int32_t p : ids) {
            int64_t o[2];
            int32_t l[2];
            AGX_CUDA(cudaMemcpyAsync(o, d_off + 2 * (int64_t)p, sizeof o, cudaMemcpyDeviceToHost, st));
            AGX_CUDA(cudaMemcpyAsync(l, d_len + 2 * (int64_t)p, sizeof l, cudaMemcpyDeviceToHost, st));
            AGX_CUDA(cudaStreamSynchronize(st));
            // columns = the longer sequence (more stripes in flight), rows = the shorter
            const int hi = l[0] >= l[1] ? 0 : 1;
            rc = sw_long_device(ws.lng, d_seqs + o[hi], l[hi], d_seqs + o[1 - hi], l[1 - hi], sc, d_scores + p, st);
            if (rc != AGX_OK) return rc;
        }
        ws.prof_long.end(st);
    }

    // generic list = pairs classified generic up front + pairs the duo kernels bounced (non-ACGT)
    const int64_t generic_upper = (int64_t)counts[GENERIC] + n_duo;
    if (generic_upper > 0 && (counts[GENERIC] > 0 || n_duo > 0)) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int64_t want_warps = generic_upper;
        const int64_t max_warps = (int64_t)sms * 16;
        if (want_warps > max_warps) want_warps = max_warps;
        const int64_t stride = 2 * (int64_t)max_len + 64;
        // keep the boundary scratch under 2 GiB: fewer (persistent) warps when rows are very long
        const int64_t fit = ((int64_t)1 << 29) / stride;
        if (want_warps > fit) want_warps = fit > WAVE_WARPS ? fit : WAVE_WARPS;
        const int blocks = (int)((want_warps + WAVE_WARPS - 1) / WAVE_WARPS);
        const int64_t need = stride * blocks * WAVE_WARPS;
        if (need > ws.cap_wave) {
            if (ws.wave_scratch) cudaFree(ws.wave_scratch);
            ws.wave_scratch = nullptr;
            ws.cap_wave = 0;
            AGX_CUDA(cudaMalloc(&ws.wave_scratch, (size_t)need * sizeof(int32_t)));
            ws.cap_wave = need;
        }
        ws.prof_wave.begin(st);
        // the coded pass needs the two substitution scores (minus goe) to fit a signed byte
        const bool coded_ok = (sc.match - goe) <= 127 && (sc.match - goe) >= -128 && (sc.mismatch - goe) <= 127 &&
                              (sc.mismatch - goe) >= -128 && getenv("AGX_WAVE_RAW") == nullptr;
        sw_wave_kernel<0><<<blocks, WAVE_WARPS * 32, 0, st>>>(
            d_seqs, d_off, d_len, ws.order + (int64_t)GENERIC * n_pairs, ws.counters + GENERIC, sc, 1,
            coded_ok ? 1 : 0, d_scores, ws.wave_scratch, stride, WaveAlignOut{});
        ws.prof_wave.end(st);
        count_launch();
        AGX_CUDA(cudaGetLastError());
    }
    return AGX_OK;
}

// ------------------------------------------------------------------------------------------
// host side of the alignment path (end cell / start cell / CIGAR)
// ------------------------------------------------------------------------------------------
namespace {
template <typename T> int grow(T *&p, int64_t &cap, int64_t need)
{
    if (need <= cap) return AGX_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const int64_t want = need + need / 8 + 64;
    cudaError_t e = cudaMalloc(&p, (size_t)want * sizeof(T));
    if (e != cudaSuccess) return fail(AGX_ENOMEM, std::string("cudaMalloc (alignment scratch): ") + cudaGetErrorString(e));
    cap = want;
    return AGX_OK;
}
}  // namespace

void sw_align_workspace_free(SwAlignWorkspace &ws)
{
    for (void *p : {(void *)ws.order, (void *)ws.cap32, (void *)ws.tmp_off, (void *)ws.cig_off, (void *)ws.gen_units,
                    (void *)ws.gen_off, (void *)ws.scan_tmp, (void *)ws.d_total, (void *)ws.wk, (void *)ws.counters,
                    (void *)ws.tb, (void *)ws.tb_gen, (void *)ws.tmp_ops, (void *)ws.wave_scratch, (void *)ws.walk_bins})
        if (p) cudaFree(p);
    if (ws.h_total) cudaFreeHost(ws.h_total);
    if (ws.h_counters) cudaFreeHost(ws.h_counters);
    ws.prof_dp.destroy();
    ws.prof_walk.destroy();
    ws = SwAlignWorkspace();
}

// bytes per ROW of the larger of the two layouts a pair can end up in: its duo class (half a duo: K2 words x 2
// bytes of the pair per lane, G lanes) or the wavefront layout (256-byte stripes); a chunk of pairs then takes
// at most (sum of these) x (its longest line + 32) bytes.  Tabulated by the shorter length for the lengths the duo
// classes cover (*max_len = the last tabulated one); longer: 256-byte stripes.
const int32_t *sw_align_tb_row_table(int32_t *max_len)
{
    static const std::vector<int32_t> table = [] {
        std::vector<int32_t> t(DUO_MAX_CAP + 2, 0);
        for (int ra = 0; ra <= DUO_MAX_CAP + 1; ++ra) {
            int32_t duo = 0;
            for (int c = 0; c < SW_N_DUO_CLASSES; ++c)
                if (ra <= duo_cap(c) + 1) { duo = duo_class(c).g * (((duo_class(c).k + 1) / 2) * 2 + 1); break; }   // + the block padding
            const int32_t wave = ((ra + 255) / 256) * 256;
            t[ra] = duo > wave ? duo : wave;
        }
        return t;
    }();
    *max_len = DUO_MAX_CAP + 1;
    return table.data();
}

int64_t sw_align_tb_row_bytes(int32_t len_a, int32_t len_b)
{
    int32_t cap = 0;
    const int32_t *table = sw_align_tb_row_table(&cap);
    const int32_t ra = len_a > len_b ? len_b : len_a;
    if (ra <= cap) return table[ra];
    return (int64_t)((ra + 255) / 256) * 256;
}

int sw_align_run_device(SwAlignWorkspace &ws, const uint8_t *d_seqs, const int64_t *d_off, const int32_t *d_len,
                        int64_t n_pairs, SwScoring sc, int mode, int64_t tb_budget, int32_t *d_scores, int32_t *d_ends,
                        int32_t *d_coords, int64_t *cigar_total, cudaStream_t st)
{
    if (cigar_total) *cigar_total = 0;
    if (n_pairs == 0) return AGX_OK;
    if (n_pairs > (int64_t)1 << 30) return fail(AGX_ERANGE, "sw align: more than 2^30 pairs in one call");
    const bool trace = getenv("AGX_ALIGN_TRACE") != nullptr;       // host-side timeline on stderr
    const auto t_begin = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (trace) fprintf(stderr, "[agx align run] +%.3f ms: %s\n",
                           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(), what);
    };
    if (!(sc.match > 0 && sc.mismatch < 0 && sc.gap_open <= 0 && sc.gap_extend < 0))
        return fail(AGX_ERANGE, "sw: scoring must satisfy match > 0 > mismatch, gap_open <= 0, gap_extend < 0");
    const int32_t goe = sc.gap_open + sc.gap_extend;
    // neighbouring H values must differ by less than 128 (the walk rebuilds H from byte differences) and the
    // keys hold 21 bits of score
    if (sc.match - goe > 127 || -sc.mismatch > 127 || -goe > 127)
        return fail(AGX_ERANGE, "sw align: match - (gap_open + gap_extend), -mismatch and -(gap_open + gap_extend) must be <= 127");

    int rc;
    if (n_pairs > ws.cap_pairs) {
        int64_t c;
        const int64_t n = n_pairs + n_pairs / 8 + 64;
#define AGX_REGROW(ptr, T, elems) c = 0; if (ptr) { cudaFree(ptr); ptr = nullptr; } if ((rc = grow(ptr, c, (int64_t)(elems))) != AGX_OK) return rc;
        ws.cap_pairs = 0;
        AGX_REGROW(ws.order, int32_t, n * SW_N_CLASSES)
        AGX_REGROW(ws.cap32, int32_t, n)
        AGX_REGROW(ws.tmp_off, int64_t, n + 1)
        AGX_REGROW(ws.cig_off, int64_t, n + 1)
        AGX_REGROW(ws.gen_units, int32_t, n)
        AGX_REGROW(ws.gen_off, int64_t, n + 1)
        AGX_REGROW(ws.scan_tmp, int64_t, device_scan_tmp_elems(n) + 8)
        AGX_REGROW(ws.wk, SwWalkRec, n)
#undef AGX_REGROW
        ws.cap_pairs = n;
    }
    if (!ws.counters) {
        AGX_CUDA(cudaMalloc(&ws.counters, CNT_A_WORDS * sizeof(int32_t)));
        AGX_CUDA(cudaMallocHost(&ws.h_counters, CNT_A_WORDS * sizeof(int32_t)));
        AGX_CUDA(cudaMalloc(&ws.d_total, 4 * sizeof(int64_t)));
        AGX_CUDA(cudaMalloc(&ws.walk_bins, 2 * WALK_BINS * sizeof(int32_t)));
        AGX_CUDA(cudaMallocHost(&ws.h_total, 4 * sizeof(int64_t)));
    }

    // s16x2 path: as in sw_run_device, and every H must fit the 10 score bits of the row keys
    const bool s16_ok = (sc.match - goe) <= 127 && (sc.mismatch - goe) >= -128 && goe >= -1024 && sc.match <= 30;
    const int32_t s16_max_short = s16_ok ? min(DUO_MAX_CAP, 1023 / sc.match) : 0;

    AGX_CUDA(cudaMemsetAsync(ws.counters, 0, CNT_A_WORDS * sizeof(int32_t), st));
    const int cblocks = (int)((n_pairs + 255) / 256);
    sw_align_classify_kernel<<<cblocks, 256, 0, st>>>(d_seqs, d_off, d_len, n_pairs, s16_max_short, sw_long_cells(), sc.match,
                                                      mode, ws.order, ws.counters, d_scores, d_ends, ws.cap32, ws.wk);
    count_launch();
    AGX_CUDA(cudaGetLastError());
    AGX_CUDA(cudaMemcpyAsync(ws.h_counters, ws.counters, CNT_A_WORDS * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    AGX_CUDA(cudaStreamSynchronize(st));
    mark("length classes read back");
    if (ws.h_counters[CNT_A_BAD] > 0)
        return fail(AGX_ERANGE, "sw align: " + std::to_string(ws.h_counters[CNT_A_BAD]) +
                                    " pair(s) outside the supported range (whole-GPU pairs of >= 2^28 cells, a shorter "
                                    "line above 16000 symbols, or lengths adding up to 2^29)");
    int32_t counts[SW_N_CLASSES], rows[SW_N_DUO_CLASSES];
    int64_t n_duo = 0, tb_base[SW_N_DUO_CLASSES], duo_total = 0;
    for (int c = 0; c < SW_N_CLASSES; ++c) counts[c] = ws.h_counters[c];
    for (int c = 0; c < SW_N_DUO_CLASSES; ++c) {
        n_duo += counts[c];
        rows[c] = ws.h_counters[CNT_A_ROWS + c];
        tb_base[c] = duo_total;
        if (mode == 2) duo_total += duo_tb_class_bytes(c, counts[c], rows[c]);
    }
    const int32_t max_len = ws.h_counters[CNT_MAXLEN];
    if (mode == 2) {
        if (duo_total > tb_budget) return fail(AGX_ENOMEM, "sw align: traceback matrices exceed the budget");
        if ((rc = grow(ws.tb, ws.cap_tb, duo_total + 256)) != AGX_OK) return rc;
    }

    DuoConst kc;
    kc.goe = goe;
    kc.match = sc.match;
    kc.goe2 = ((uint32_t)(uint16_t)(int16_t)goe) * 0x00010001u;
    kc.ext2 = ((uint32_t)(uint16_t)(int16_t)sc.gap_extend) * 0x00010001u;
    const uint32_t xb = (uint32_t)(uint8_t)(int8_t)(sc.mismatch - goe);
    const uint32_t mb = (uint32_t)(uint8_t)(int8_t)(sc.match - goe);
    kc.xb4 = xb * 0x01010101u;
    kc.mxor = xb ^ mb;

    DuoAlignOut ao = {};
    ao.ends = d_ends;
    ao.wk = ws.wk;
    ao.tb = mode == 2 ? ws.tb : nullptr;
    ao.k32 = 32;
    ws.prof_dp.begin(st);
    rc = launch_duo_align<0>(d_seqs, d_off, d_len, ws.order, n_pairs, counts, rows, tb_base, mode, kc, d_scores,
                             ws.counters, ao, st);
    if (rc != AGX_OK) return rc;

    // wavefront pairs: those classified so + those the duo kernels bounced (bytes outside ACGT)
    const int64_t generic_upper = (int64_t)counts[GENERIC] + n_duo;
    if (generic_upper > 0) {
        const int32_t *glist = ws.order + (int64_t)GENERIC * n_pairs;
        if (mode == 2) {
            const int gb = (int)((generic_upper + 255) / 256);
            sw_wave_tb_units_kernel<<<gb, 256, 0, st>>>(d_len, glist, ws.counters + GENERIC, ws.gen_units, generic_upper);
            count_launch();
            if ((rc = device_exclusive_scan(ws.gen_units, generic_upper, ws.gen_off, ws.scan_tmp, ws.d_total, st)) != AGX_OK) return rc;
            AGX_CUDA(cudaMemcpyAsync(ws.h_total, ws.d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
            AGX_CUDA(cudaStreamSynchronize(st));
            const int64_t gen_bytes = ws.h_total[0] * 256;
            if (duo_total + gen_bytes > tb_budget) return fail(AGX_ENOMEM, "sw align: traceback matrices exceed the budget");
            if ((rc = grow(ws.tb_gen, ws.cap_tb_gen, gen_bytes + 256)) != AGX_OK) return rc;
            units_to_bytes_kernel<<<gb, 256, 0, st>>>(ws.gen_off, generic_upper, 0);
            count_launch();
        }
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int64_t want_warps = generic_upper;
        const int64_t max_warps = (int64_t)sms * 16;
        if (want_warps > max_warps) want_warps = max_warps;
        const int64_t stride = 2 * (int64_t)max_len + 64;
        const int64_t fit = ((int64_t)1 << 29) / stride;
        if (want_warps > fit) want_warps = fit > WAVE_WARPS ? fit : WAVE_WARPS;
        const int blocks = (int)((want_warps + WAVE_WARPS - 1) / WAVE_WARPS);
        if ((rc = grow(ws.wave_scratch, ws.cap_wave, stride * blocks * WAVE_WARPS)) != AGX_OK) return rc;
        const bool coded_ok = (sc.match - goe) <= 127 && (sc.match - goe) >= -128 && (sc.mismatch - goe) <= 127 &&
                              (sc.mismatch - goe) >= -128 && getenv("AGX_WAVE_RAW") == nullptr;
        WaveAlignOut wo = {};
        wo.ends = d_ends;
        wo.wk = ws.wk;
        wo.tb = ws.tb_gen;
        wo.tb_off = ws.gen_off;
        wo.k32 = 32;
        if (mode == 2)
            sw_wave_kernel<2><<<blocks, WAVE_WARPS * 32, 0, st>>>(d_seqs, d_off, d_len, glist, ws.counters + GENERIC, sc, 1,
                                                               coded_ok ? 1 : 0, d_scores, ws.wave_scratch, stride, wo);
        else
            sw_wave_kernel<1><<<blocks, WAVE_WARPS * 32, 0, st>>>(d_seqs, d_off, d_len, glist, ws.counters + GENERIC, sc, 1,
                                                               coded_ok ? 1 : 0, d_scores, ws.wave_scratch, stride, wo);
        count_launch();
        AGX_CUDA(cudaGetLastError());
    }
    ws.prof_dp.end(st);
    mark("DP kernels queued");
    if (mode != 2) return AGX_OK;

    // the walk: room for the runs of every pair, then one thread per pair
    if ((rc = device_exclusive_scan(ws.cap32, n_pairs, ws.tmp_off, ws.scan_tmp, ws.d_total + 1, st)) != AGX_OK) return rc;
    AGX_CUDA(cudaMemcpyAsync(ws.h_total + 1, ws.d_total + 1, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    AGX_CUDA(cudaStreamSynchronize(st));
    mark("DP kernels done, run capacities scanned");
    if ((rc = grow(ws.tmp_ops, ws.cap_tmp, ws.h_total[1] + 16)) != AGX_OK) return rc;
    AGX_CUDA(cudaMemsetAsync(ws.d_total + 2, 0, sizeof(int64_t), st));
    ws.prof_walk.begin(st);
    // walks of similar length side by side: pairs ordered by the size of their score (ws.gen_units is free again)
    int32_t *walk_bins = ws.walk_bins, *walk_perm = ws.gen_units;
    AGX_CUDA(cudaMemsetAsync(walk_bins, 0, 2 * WALK_BINS * sizeof(int32_t), st));
    sw_walk_count_kernel<<<cblocks, 256, 0, st>>>(d_scores, n_pairs, walk_bins);
    sw_walk_order_kernel<<<cblocks, 256, 0, st>>>(d_scores, n_pairs, walk_bins, walk_perm);
    count_launch(2);
    // the two matrix buffers are addressed from one base: wavefront records carry an offset relative to tb_gen
    sw_walk_kernel<<<(int)((n_pairs + 127) / 128), 128, 0, st>>>(d_seqs, d_off, d_len, n_pairs, sc, ws.wk, ws.tb, ws.tb_gen, d_scores,
                                                                d_ends, d_coords, ws.tmp_ops, ws.tmp_off, ws.cap32,
                                                                reinterpret_cast<int32_t *>(ws.d_total + 2), walk_perm);
    count_launch();
    ws.prof_walk.end(st);
    AGX_CUDA(cudaGetLastError());
    if ((rc = device_exclusive_scan(ws.cap32, n_pairs, ws.cig_off, ws.scan_tmp, ws.cig_off + n_pairs, st)) != AGX_OK) return rc;
    AGX_CUDA(cudaMemcpyAsync(ws.h_total + 2, ws.d_total + 2, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    AGX_CUDA(cudaMemcpyAsync(ws.h_total + 3, ws.cig_off + n_pairs, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    AGX_CUDA(cudaStreamSynchronize(st));
    mark("walk done, runs scanned");
    if (ws.h_total[2] != 0) return fail(AGX_ECUDA, "sw align: the traceback walk lost its path (internal error)");
    if (cigar_total) *cigar_total = ws.h_total[3];
    return AGX_OK;
}

int sw_align_gather_device(SwAlignWorkspace &ws, int64_t n_pairs, uint32_t *d_cigar, cudaStream_t st)
{
    if (n_pairs == 0) return AGX_OK;
    sw_cigar_gather_kernel<<<(int)((n_pairs + 255) / 256), 256, 0, st>>>(ws.tmp_ops, ws.tmp_off, ws.cap32, ws.cig_off, n_pairs, d_cigar);
    count_launch();
    AGX_CUDA(cudaGetLastError());
    return AGX_OK;
}

}  // namespace agx
