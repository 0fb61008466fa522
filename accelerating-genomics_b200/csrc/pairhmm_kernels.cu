// pairhmm_kernels.cu -- PairHMM forward (M / insertion X / deletion Y) on B200 (sm_100a).
//
// Replaces pairHMM/antidiagsPairHMM.c:99-117 (prior setup), :120-241 (recurrence), :206-212 and
// :242 (final log10 sum); the same maths is pairHMM/pairHMMmatrix.c:20-66.  Restated
// (SURVEY.md section 8a, rows a6-a10), with Q* = 10^-((c-33)/10) of the read row i-1:
//     M[i][j] = prior(i,j) * ((1-(Qi+Qd)) * M[i-1][j-1] + (1-Qg) * (X[i-1][j-1] + Y[i-1][j-1]))
//     X[i][j] = M[i-1][j] * Qi + X[i-1][j] * Qg
//     Y[i][j] = M[i][j-1] * Qd + Y[i][j-1] * Qg
//     prior   = (r == h || r == 'N' || h == 'N') ? 1-Qr : Qr          (reference quirk: no /3)
//     row 0: M = X = 0, Y = SCALE / hap_len;  column 0 (i >= 1): 0
//     result  = log10(sum_j M[R][j] + X[R][j]) - log10(SCALE)
//
// Three kernels:
//   hmm_duo_kernel<K>      FP32 fast path, TWO reads of the same batch and row class per warp: the same
//                          streaming scheme as hmm_stream_kernel below, with read A in the low and read B
//                          in the high half of every value and every FP32 instruction in its packed
//                          sm_100 form (FFMA2 / FMUL2 / FADD2 = fma.rn.f32x2 ...): one issue slot per two
//                          cells, and the per-step costs (three boundary shuffles, prior lookup, symbol
//                          prefetch, haplotype counter) are paid once for 2*K cells instead of K.
//   hmm_stream_kernel<K>   FP32 fast path for the reads left without a partner.  One warp per read; lane t owns K consecutive read rows
//                          in registers (rows are bottom-aligned, so row R is always the last row
//                          of lane 31); ALL haplotypes of the read's batch stream through the warp
//                          back to back, column by column, lane t one column behind lane t-1
//                          (the bottom row of a lane moves down with __shfl_up_sync).  The
//                          recurrence is carried on X' = X/Qi_i and Y' = Y/Qd_i, which turns the
//                          X and Y updates into one FMA each (6 FP32 instructions per cell instead
//                          of 8).  SCALE = 2^120.  Sums that are not finite or are too small to
//                          be trusted in FP32 are queued for the FP64 kernel.
//   hmm_striped_kernel<T,K> generic path: one warp per (read, haplotype), any read length (row
//                          stripes of 32*K rows chained through a boundary row in global memory).
//                          With T = double it evaluates the reference's expression in the
//                          reference's own association order without FMA contraction and with
//                          SCALE = DBL_MAX/16, i.e. it is the FP64 rescue / exact-parity path.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace agx {

namespace {

constexpr int HMM_MAX_K = 8;                    // stream kernel: reads up to 256 rows
constexpr int HMM_N_CLASSES = HMM_MAX_K + 1;    // classes 0..7 -> K = 1..8, class 8 -> striped
constexpr int HMM_LONG = HMM_MAX_K;
constexpr int HCNT_RESCUE = HMM_N_CLASSES;      // counters[HCNT_RESCUE]  = rescue list length
constexpr int HCNT_MAXHAP = HMM_N_CLASSES + 1;  // counters[HCNT_MAXHAP]  = longest haplotype
constexpr int HCNT_UNSORTED = HMM_N_CLASSES + 2;   // != 0: read_batch is not non-decreasing (no pairing)
constexpr int HCNT_PAIRS = HMM_N_CLASSES + 4;   // counters[HCNT_PAIRS + c] = read pairs of row class c
constexpr int HCNT_WORDS = HCNT_PAIRS + HMM_MAX_K;
#ifndef AGX_HMM_WARPS
#define AGX_HMM_WARPS 1
#endif
#ifndef AGX_HMM_MINBLOCKS
#define AGX_HMM_MINBLOCKS 20
#endif
constexpr int HMM_WARPS = AGX_HMM_WARPS;
// measured on BASELINE configs[3] (profiles/r2ag_hmm_variants.jsonl, r2ah_hmm_variants.jsonl; kernel ms per 10^6 pairs):
// 4 -> 21.39, 8 -> 21.07, 12 -> 20.70, 16 -> 20.62, 24 -> 20.76, 32 -> 21.44
#ifndef AGX_HMM_FAST
#define AGX_HMM_FAST 16
#endif
constexpr int HMM_FAST = AGX_HMM_FAST;           // branch-free column steps per loop trip of hmm_duo_kernel

constexpr float SCALE_F = 1.329227995784916e36f;  // 2^120
// a forward sum below this (2^-100) may have lost low-order terms to FP32 underflow
constexpr float RESCUE_BELOW = 7.888609052210118e-31f;

__global__ void __launch_bounds__(256)
hmm_classify_kernel(const int32_t *__restrict__ read_len, int64_t n_reads,
                    const int32_t *__restrict__ hap_len, int64_t n_haps,
                    int32_t *__restrict__ order, int32_t *__restrict__ counters, int force_long)
{
    __shared__ int32_t s_cnt[HMM_N_CLASSES];
    __shared__ int32_t s_base[HMM_N_CLASSES];
    __shared__ int32_t s_maxhap;
    if (threadIdx.x < HMM_N_CLASSES) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_maxhap = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int cls = -1, rank = 0;
    if (i < n_reads) {
        const int32_t R = read_len[i];
        cls = (R <= 32 * HMM_MAX_K && !force_long) ? (R + 31) / 32 - 1 : HMM_LONG;
        if (cls < 0) cls = 0;
        rank = atomicAdd(&s_cnt[cls], 1);
    }
    if (i < n_haps) atomicMax(&s_maxhap, hap_len[i]);
    __syncthreads();
    if (threadIdx.x < HMM_N_CLASSES && s_cnt[threadIdx.x] > 0)
        s_base[threadIdx.x] = atomicAdd(&counters[threadIdx.x], s_cnt[threadIdx.x]);
    if (threadIdx.x == 0 && s_maxhap > 0) atomicMax(&counters[HCNT_MAXHAP], s_maxhap);
    __syncthreads();
    if (cls >= 0) order[(int64_t)cls * n_reads + s_base[cls] + rank] = (int32_t)i;
}

// ------------------------------------------------------------------------------------------
// read pairing for hmm_duo_kernel
// ------------------------------------------------------------------------------------------
// first / one-past-last read of every batch (reads of a batch are consecutive when read_batch is
// non-decreasing; anything else raises HCNT_UNSORTED and the host falls back to one read per warp)
__global__ void __launch_bounds__(256)
hmm_ranges_kernel(const int32_t *__restrict__ read_batch, int64_t n_reads, const int32_t *__restrict__ hap_len,
                  int64_t n_haps, int32_t *__restrict__ batch_first, int32_t *__restrict__ batch_end,
                  int32_t *__restrict__ counters)
{
    __shared__ int32_t s_maxhap;
    if (threadIdx.x == 0) s_maxhap = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_reads) {
        const int32_t b = read_batch[i];
        if (i == 0 || read_batch[i - 1] != b) batch_first[b] = (int32_t)i;
        if (i == n_reads - 1 || read_batch[i + 1] != b) batch_end[b] = (int32_t)i + 1;
        if (i > 0 && read_batch[i - 1] > b) counters[HCNT_UNSORTED] = 1;
    }
    if (i < n_haps) atomicMax(&s_maxhap, hap_len[i]);
    __syncthreads();
    if (threadIdx.x == 0 && s_maxhap > 0) atomicMax(&counters[HCNT_MAXHAP], s_maxhap);
}

// One warp per batch: reads of the same row class are paired in file order; a class with an odd number
// of reads leaves one single (-> hmm_stream_kernel); reads longer than the stream kernels take go to the
// striped kernel's list.  pairs[c][...] / order[c][...] are filled through warp-aggregated atomics.
__global__ void __launch_bounds__(128)
hmm_pair_kernel(const int32_t *__restrict__ read_len, int64_t n_reads, int64_t n_batches,
                const int32_t *__restrict__ batch_first, const int32_t *__restrict__ batch_end,
                int2 *__restrict__ pairs, int32_t *__restrict__ order, int32_t *__restrict__ counters)
{
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp; b < n_batches; b += n_warps) {
        const int32_t rs = batch_first[b], re = batch_end[b];
        if (rs < 0 || re <= rs) continue;
        int32_t pend[HMM_MAX_K];
#pragma unroll
        for (int c = 0; c < HMM_MAX_K; ++c) pend[c] = -1;
        for (int32_t base = rs; base < re; base += 32) {
            const int32_t r = base + lane;
            int cls = -1;
            if (r < re) {
                const int32_t R = read_len[r];
                cls = R <= 32 * HMM_MAX_K ? (R + 31) / 32 - 1 : HMM_LONG;
                if (cls < 0) cls = 0;
            }
            const uint32_t mlong = __ballot_sync(0xffffffffu, cls == HMM_LONG);
            if (mlong) {
                int32_t pos = 0;
                if (lane == 0) pos = atomicAdd(&counters[HMM_LONG], __popc(mlong));
                pos = __shfl_sync(0xffffffffu, pos, 0);
                if (cls == HMM_LONG) order[(int64_t)HMM_LONG * n_reads + pos + __popc(mlong & lt)] = r;
            }
#pragma unroll
            for (int c = 0; c < HMM_MAX_K; ++c) {
                const uint32_t m = __ballot_sync(0xffffffffu, cls == c);
                if (m == 0) continue;
                const int hasp = pend[c] >= 0 ? 1 : 0;
                const int total = __popc(m) + hasp;
                const int np = total >> 1;
                int32_t pos = 0;
                if (np > 0 && lane == 0) pos = atomicAdd(&counters[HCNT_PAIRS + c], np);
                pos = __shfl_sync(0xffffffffu, pos, 0);
                const int e = __popc(m & lt) + hasp;          // my index in [pending read] + this tile's reads
                if (cls == c && (e & 1)) {
                    // second read of a pair: the first is the pending read or the lane just before me in m
                    const int32_t first = (e == 1 && hasp) ? pend[c] : base + (int32_t)__fns(m, 0, e - hasp);
                    pairs[(int64_t)c * n_reads + pos + (e >> 1)] = make_int2(first, r);
                }
                // an odd total leaves the tile's last read of this class pending
                pend[c] = (total & 1) ? base + (31 - __clz(m)) : -1;
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < HMM_MAX_K; ++c)
                if (pend[c] >= 0) order[(int64_t)c * n_reads + atomicAdd(&counters[c], 1)] = pend[c];
        }
    }
}

// ------------------------------------------------------------------------------------------
// haplotype preparation: symbol codes + per-haplotype record
// ------------------------------------------------------------------------------------------
// Haplotype symbols are recoded once per call to the byte offset of their row in the per-warp prior
// table of the stream kernel: A 0, C 1, T 2, G 3, N 4 (the wildcard), anything else 5.  A haplotype
// holding a symbol outside "ACGTN" cannot use the table (the reference compares raw bytes, so such a
// symbol may still equal a read base); its record carries init = NaN, which makes every forward sum
// NaN and sends its pairs to the byte-exact FP64 kernel.
struct __align__(16) HapInfo {
    int64_t off;     // offset of the haplotype in buf (and of its codes in the code buffer)
    int32_t len;
    float init;      // SCALE / len, or NaN
};

constexpr int HMM_NSYM = 6;

// base-quality character as the prior setup reads it: GATK mode with the floor reads qualities below 6 as 6
__device__ __forceinline__ uint32_t base_qual(uint32_t ch, int gatk)
{
    return ((gatk & 2) && (int)(signed char)ch < 33 + 6) ? 33u + 6u : ch;
}

__device__ __forceinline__ uint32_t hap_symbol(uint32_t ch)
{
    const uint32_t code = (ch >> 1) & 3u;                       // A 0, C 1, T 2, G 3
    if (((0x47544341u >> (8 * code)) & 0xffu) == ch) return code;
    return ch == 'N' ? 4u : 5u;
}

__global__ void __launch_bounds__(128)
hmm_prep_haps_kernel(const uint8_t *__restrict__ buf, const int64_t *__restrict__ hap_off,
                     const int32_t *__restrict__ hap_len, int64_t n_haps, uint8_t *__restrict__ codes,
                     HapInfo *__restrict__ info)
{
    __shared__ int s_bad;
    for (int64_t h = blockIdx.x; h < n_haps; h += gridDim.x) {
        if (threadIdx.x == 0) s_bad = 0;
        __syncthreads();
        const int64_t off = hap_off[h];
        const int32_t len = hap_len[h];
        int bad = 0;
        for (int32_t i = threadIdx.x; i < len; i += blockDim.x) {
            const uint32_t sym = hap_symbol(buf[off + i]);
            bad |= (sym == 5u);
            codes[off + i] = (uint8_t)sym;
        }
        if (bad) s_bad = 1;
        __syncthreads();
        if (threadIdx.x == 0) {
            HapInfo hi;
            hi.off = off;
            hi.len = len;
            hi.init = s_bad ? nanf("") : SCALE_F / (float)len;
            info[h] = hi;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// FP32 streaming kernel
// ------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(HMM_WARPS * 32, AGX_HMM_MINBLOCKS)
hmm_stream_kernel(HmmBatchView v, const int64_t *__restrict__ read_out_off,
                  const int32_t *__restrict__ order_cls, int32_t n_items,
                  const double *__restrict__ lut_g, int gatk, const uint8_t *__restrict__ codes,
                  const uint8_t *__restrict__ zero_pad, const HapInfo *__restrict__ hapinfo,
                  float *__restrict__ sums)
{
    constexpr int CH = (K + 3) / 4;                    // float4 chunks of priors per lane
    __shared__ double lut[256];
    __shared__ float4 prior_tab[HMM_WARPS][HMM_NSYM][CH][32];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = lut_g[i];
    __syncthreads();

    const int t = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int item = blockIdx.x * HMM_WARPS + wib;
    if (item >= n_items) return;
    const int32_t r = order_cls[item];
    const int32_t R = v.read_len[r];
    const int32_t bt = v.read_batch[r];
    const int32_t h0 = (int32_t)v.batch_hap_start[bt], h1 = (int32_t)v.batch_hap_start[bt + 1];
    if (h1 <= h0) return;
    float *my_sums = sums + read_out_off[r];

    // ---- prior / transition setup for this lane's K rows (rows bottom-aligned) ----------------
    const int pad = 32 * K - R;
    const uint8_t *f_b = v.buf + v.read_field_off[5 * (int64_t)r + 0];
    const uint8_t *f_q = v.buf + v.read_field_off[5 * (int64_t)r + 1];
    const uint8_t *f_i = v.buf + v.read_field_off[5 * (int64_t)r + 2];
    const uint8_t *f_d = v.buf + v.read_field_off[5 * (int64_t)r + 3];
    const uint8_t *f_g = v.buf + v.read_field_off[5 * (int64_t)r + 4];

    float ca[K], cbx[K], cby[K], ccx[K], cg[K];
    {
        float pm[CH * 4], px[CH * 4];
        uint32_t rsym[CH * 4];
#pragma unroll
        for (int jj = 0; jj < CH * 4; ++jj) { pm[jj] = px[jj] = 0.f; rsym[jj] = 5u; }
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
            const int i = t * K + jj - pad;      // 0-based read position of this row
            if (i < 0) {
                // padding rows reproduce row 0: M = X = 0 and Y stays at its initial value
                ca[jj] = cbx[jj] = cby[jj] = ccx[jj] = 0.f; cg[jj] = 1.f;
            } else {
                const double Qr = lut[base_qual(f_q[i], gatk)], Qi = lut[f_i[i]], Qd = lut[f_d[i]], Qg = lut[f_g[i]];
                // previous row's Qi / Qd un-scale X' and Y' of the diagonal cell; row 0 is unscaled
                const double Qi_up = (i > 0) ? lut[f_i[i - 1]] : 1.0;
                const double Qd_up = (i > 0) ? lut[f_d[i - 1]] : 1.0;
                const uint32_t base = f_b[i];
                const double mm = 1.0 - (Qi + Qd), gm = 1.0 - Qg;
                pm[jj] = (float)(1.0 - Qr);
                px[jj] = (float)((gatk & 1) ? Qr / 3.0 : Qr);
                rsym[jj] = (base == 'N') ? 4u : hap_symbol(base);   // 5: matches nothing but N
                ca[jj] = (float)mm;
                cbx[jj] = (float)(gm * Qi_up);
                cby[jj] = (float)(gm * Qd_up);
                ccx[jj] = (i > 0) ? (float)(Qg * Qi_up / Qi) : 0.f;
                cg[jj] = (float)Qg;
            }
        }
        // priors per haplotype symbol: p(r, h) = (r == h || r == 'N' || h == 'N') ? 1-Qr : Qr
#pragma unroll
        for (int sym = 0; sym < HMM_NSYM; ++sym) {
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                float q[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int jj = ch * 4 + e;
                    const bool match = (sym == 4) || (rsym[jj] == 4u) || (sym < 4 && rsym[jj] == (uint32_t)sym);
                    q[e] = match ? pm[jj] : px[jj];
                }
                prior_tab[wib][sym][ch][t] = make_float4(q[0], q[1], q[2], q[3]);
            }
        }
    }
    __syncwarp();
    const float qi_last = (float)lut[f_i[R - 1]];   // X[R][j] = Qi_R * X'[R][j]
    const bool top_boundary = (t * K - 1) < pad;    // the row above this lane's first row is row 0
    const int n_pad_rows = pad - t * K;             // rows jj < n_pad_rows of this lane are padding

    // ---- streaming state ------------------------------------------------------------------------
    float M[K], X[K], Y[K];
#pragma unroll
    for (int jj = 0; jj < K; ++jj) { M[jj] = X[jj] = Y[jj] = 0.f; }
    float pdM = 0.f, pdX = 0.f, pdY = 0.f;           // what arrived one step ago = diagonal inputs
    float bM = 0.f, bX = 0.f, bY = 0.f;              // this lane's bottom row, sent down next step
    float acc = 0.f, init = 0.f;

    // Per-lane haplotype cursor.  Lane t starts on a virtual haplotype of t columns (it idles for t
    // steps) whose symbols come from a zeroed pad; `rem` counts the columns left in the current
    // haplotype and `cp` walks its symbol codes (stride `inc`, 0 once the lane has drained).
    int32_t hidx = h0 - 1;
    int32_t rem = t;
    const uint8_t *cp = zero_pad;
    int64_t inc = 0;
    uint32_t code_next = 0u;

    int32_t total = 31;
    for (int32_t h = h0; h < h1; ++h) total += hapinfo[h].len;

    const float4 *tab_lane = &prior_tab[wib][0][0][t];
    constexpr int SYM_STRIDE = CH * 32;              // float4 elements between symbols
    const bool lane0 = (t == 0);

#pragma unroll 2
    for (int32_t s = 0; s < total; ++s) {
        if (rem == 0) {
            // ---- this lane finished a haplotype: lane 31 owns row R and emits the forward sum ----
            if (t == 31 && hidx >= h0) my_sums[hidx - h0] = acc;
            ++hidx;
            acc = 0.f;
            if (hidx < h1) {
                const HapInfo hi = hapinfo[hidx];
                rem = hi.len;
                cp = codes + hi.off;
                inc = 1;
                init = hi.init;
                code_next = *cp;
            } else {
                rem = 0x7fffffff;    // drained: keep stepping on neutral input
                cp = zero_pad;
                inc = 0;
                code_next = 0u;
            }
#pragma unroll
            for (int jj = 0; jj < K; ++jj) { M[jj] = 0.f; X[jj] = 0.f; Y[jj] = (jj < n_pad_rows) ? init : 0.f; }
            pdM = 0.f; pdX = 0.f; pdY = top_boundary ? init : 0.f;
        }
        const uint32_t code = code_next;
        cp += inc;
        code_next = *cp;                                  // prefetch the next column's symbol
        float4 pr4[CH];
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) pr4[ch] = tab_lane[code * SYM_STRIDE + ch * 32];

        float upM = __shfl_up_sync(0xffffffffu, bM, 1);
        float upX = __shfl_up_sync(0xffffffffu, bX, 1);
        float upY = __shfl_up_sync(0xffffffffu, bY, 1);
        if (lane0) { upM = 0.f; upX = 0.f; upY = init; }
        float dM = pdM, dX = pdX, dY = pdY;
        pdM = upM; pdX = upX; pdY = upY;
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
            const float oM = M[jj], oX = X[jj], oY = Y[jj];
            const float4 p4 = pr4[jj / 4];
            const float pr = (jj % 4 == 0) ? p4.x : (jj % 4 == 1) ? p4.y : (jj % 4 == 2) ? p4.z : p4.w;
            float vv = cby[jj] * dY;
            vv = fmaf(cbx[jj], dX, vv);
            vv = fmaf(ca[jj], dM, vv);
            const float mn = pr * vv;
            const float xn = fmaf(ccx[jj], upX, upM);
            const float yn = fmaf(cg[jj], oY, oM);
            dM = oM; dX = oX; dY = oY;
            upM = mn; upX = xn;
            M[jj] = mn; X[jj] = xn; Y[jj] = yn;
        }
        bM = M[K - 1]; bX = X[K - 1]; bY = Y[K - 1];
        acc += fmaf(qi_last, bX, bM);
        --rem;
    }
    // the last haplotype of lane 31 ends exactly at the last step
    if (t == 31 && hidx >= h0 && hidx < h1 && rem == 0) my_sums[hidx - h0] = acc;
}

// ------------------------------------------------------------------------------------------
// FP32 streaming kernel, two reads per warp, packed f32x2 arithmetic
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}

// Resident one-warp blocks per SM the register budget is cut for.  Each of the four SM sub-partitions has
// its own 16 K registers, so the useful targets are multiples of four: 12 blocks -> 168 registers per
// thread, 16 -> 128, 20 -> 96.
// (same sweep: the "tight" table below -- 128 registers for K = 6, 96 for K = 4 -- 21.07 -> 20.91 ms, 20.58 ms together
// with 16 steps per trip; the roomy one 21.48 ms)
#ifndef AGX_DUO_OCC
#define AGX_DUO_OCC 2
#endif
__host__ __device__ constexpr int duo_min_blocks(int K)
{
    return AGX_DUO_OCC == 1   ? (K >= 8 ? 8 : K >= 5 ? 12 : K == 4 ? 16 : 20)      // roomy
           : AGX_DUO_OCC == 2 ? (K >= 7 ? 12 : K >= 5 ? 16 : 20)                   // tight
                              : (K >= 6 ? 12 : K >= 4 ? 16 : 20);
}

template <int K>
__global__ void __launch_bounds__(32, duo_min_blocks(K))
hmm_duo_kernel(HmmBatchView v, const int64_t *__restrict__ read_out_off, const int2 *__restrict__ items,
               int32_t n_items, const double *__restrict__ lut, int gatk, const uint8_t *__restrict__ codes,
               const uint8_t *__restrict__ zero_pad, const HapInfo *__restrict__ hapinfo,
               float *__restrict__ sums)
{
    constexpr int CH = (K + 1) / 2;                    // float4 = (row 2c: A, B; row 2c+1: A, B)
    constexpr int CH4 = K / 2;                         // whole float4 chunks; an odd K leaves one row over
    // The odd row has its own float2 table: as the half-used last float4 of prior_tab the compiler fetched it with a
    // 64-bit load at a 16-byte lane stride -- a two-way bank conflict in every step (74 M / 84 M conflicts in the K = 7
    // and K = 5 launches, 9 % of their shared-memory wavefronts, against 5 M for K = 8: profiles/r2be_hmm_duo_ncu.txt)
    __shared__ float4 prior_tab[HMM_NSYM][CH4 > 0 ? CH4 : 1][32];
    __shared__ float2 prior_odd[HMM_NSYM][(K & 1) ? 32 : 1];

    const int t = threadIdx.x;
    const int item = blockIdx.x;
    if (item >= n_items) return;
    const int2 rr = items[item];
    const int32_t bt = v.read_batch[rr.x];             // both reads belong to this batch
    const int32_t h0 = (int32_t)v.batch_hap_start[bt], h1 = (int32_t)v.batch_hap_start[bt + 1];
    if (h1 <= h0) return;
    float *sums_a = sums + read_out_off[rr.x];
    float *sums_b = sums + read_out_off[rr.y];

    // ---- prior / transition setup: .x = read A, .y = read B (rows of both bottom-aligned) ---------
    float2 ca[K], cbx[K], cby[K], ccx[K], cg[K];
    float2 qi_last;
    int n_pad_rows[2];
    bool top_boundary[2];
    {
        float pm[2][CH * 2], px[2][CH * 2];
        uint32_t rsym[2][CH * 2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int32_t r = h ? rr.y : rr.x;
            const int32_t R = v.read_len[r];
            const int pad = 32 * K - R;
            const uint8_t *f_b = v.buf + v.read_field_off[5 * (int64_t)r + 0];
            const uint8_t *f_q = v.buf + v.read_field_off[5 * (int64_t)r + 1];
            const uint8_t *f_i = v.buf + v.read_field_off[5 * (int64_t)r + 2];
            const uint8_t *f_d = v.buf + v.read_field_off[5 * (int64_t)r + 3];
            const uint8_t *f_g = v.buf + v.read_field_off[5 * (int64_t)r + 4];
#pragma unroll
            for (int jj = 0; jj < CH * 2; ++jj) { pm[h][jj] = px[h][jj] = 0.f; rsym[h][jj] = 5u; }
#pragma unroll
            for (int jj = 0; jj < K; ++jj) {
                const int i = t * K + jj - pad;      // 0-based read position of this row
                float a = 0.f, bx = 0.f, by = 0.f, cx = 0.f, g = 1.f;   // padding rows reproduce row 0
                if (i >= 0) {
                    const double Qr = __ldg(lut + base_qual(f_q[i], gatk)), Qi = __ldg(lut + f_i[i]), Qd = __ldg(lut + f_d[i]),
                                 Qg = __ldg(lut + f_g[i]);
                    const double Qi_up = (i > 0) ? __ldg(lut + f_i[i - 1]) : 1.0;
                    const double Qd_up = (i > 0) ? __ldg(lut + f_d[i - 1]) : 1.0;
                    const uint32_t base = f_b[i];
                    const double mm = 1.0 - (Qi + Qd), gm = 1.0 - Qg;
                    pm[h][jj] = (float)(1.0 - Qr);
                    px[h][jj] = (float)((gatk & 1) ? Qr / 3.0 : Qr);
                    rsym[h][jj] = (base == 'N') ? 4u : hap_symbol(base);
                    a = (float)mm;
                    bx = (float)(gm * Qi_up);
                    by = (float)(gm * Qd_up);
                    cx = (i > 0) ? (float)(Qg * Qi_up / Qi) : 0.f;
                    g = (float)Qg;
                }
                if (h == 0) { ca[jj].x = a; cbx[jj].x = bx; cby[jj].x = by; ccx[jj].x = cx; cg[jj].x = g; }
                else        { ca[jj].y = a; cbx[jj].y = bx; cby[jj].y = by; ccx[jj].y = cx; cg[jj].y = g; }
            }
            const float ql = (float)__ldg(lut + f_i[R - 1]);   // X[R][j] = Qi_R * X'[R][j]
            if (h == 0) qi_last.x = ql; else qi_last.y = ql;
            top_boundary[h] = (t * K - 1) < pad;     // the row above this lane's first row is row 0
            n_pad_rows[h] = pad - t * K;             // rows jj < n_pad_rows of this lane are padding
        }
#pragma unroll
        for (int sym = 0; sym < HMM_NSYM; ++sym) {
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                float q[2][2];
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int jj = ch * 2 + e;
                        const bool match = (sym == 4) || (rsym[h][jj] == 4u) || (sym < 4 && rsym[h][jj] == (uint32_t)sym);
                        q[e][h] = match ? pm[h][jj] : px[h][jj];
                    }
                if (ch < CH4) prior_tab[sym][ch][t] = make_float4(q[0][0], q[0][1], q[1][0], q[1][1]);
                else prior_odd[sym][t] = make_float2(q[0][0], q[0][1]);
            }
        }
    }
    __syncwarp();

    // ---- streaming state ------------------------------------------------------------------------
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 M[K], X[K], Y[K];
#pragma unroll
    for (int jj = 0; jj < K; ++jj) { M[jj] = X[jj] = Y[jj] = zero2; }
    float2 pdM = zero2, pdX = zero2, pdY = zero2;    // what arrived one step ago = diagonal inputs
    float2 bM = zero2, bX = zero2, bY = zero2;       // this lane's bottom row, sent down next step
    float2 acc = zero2;
    float init = 0.f;

    int32_t hidx = h0 - 1;
    int32_t rem = t;
    const uint8_t *cp = zero_pad;
    int64_t inc = 0;
    uint32_t code_next = 0u;

    int32_t total = 31;
    for (int32_t h = h0; h < h1; ++h) total += hapinfo[h].len;

    const float4 *tab_lane = &prior_tab[0][0][t];
    const float2 *odd_lane = &prior_odd[0][(K & 1) ? t : 0];
    constexpr int SYM_STRIDE = (CH4 > 0 ? CH4 : 1) * 32;   // float4 elements between symbols
    const bool lane0 = (t == 0);

    // One column step (no haplotype switch inside: see the driver loop below).
    auto step = [&]() {
            const uint32_t code = code_next;
            cp += inc;
            code_next = *cp;                                  // prefetch the next column's symbol
            float4 pr4[CH];
#pragma unroll
            for (int ch = 0; ch < CH4; ++ch) pr4[ch] = tab_lane[code * SYM_STRIDE + ch * 32];
            if (K & 1) {
                const float2 po = odd_lane[code * 32];
                pr4[CH - 1] = make_float4(po.x, po.y, 0.f, 0.f);
            }

            float2 upM, upX, upY;
            upM.x = __shfl_up_sync(0xffffffffu, bM.x, 1); upM.y = __shfl_up_sync(0xffffffffu, bM.y, 1);
            upX.x = __shfl_up_sync(0xffffffffu, bX.x, 1); upX.y = __shfl_up_sync(0xffffffffu, bX.y, 1);
            upY.x = __shfl_up_sync(0xffffffffu, bY.x, 1); upY.y = __shfl_up_sync(0xffffffffu, bY.y, 1);
            if (lane0) { upM = zero2; upX = zero2; upY = make_float2(init, init); }
            float2 dM = pdM, dX = pdX, dY = pdY;
            pdM = upM; pdX = upX; pdY = upY;
#pragma unroll
            for (int jj = 0; jj < K; ++jj) {
                const float2 oM = M[jj], oX = X[jj], oY = Y[jj];
                const float4 p4 = pr4[jj / 2];
                const float2 pr = (jj & 1) ? make_float2(p4.z, p4.w) : make_float2(p4.x, p4.y);
                float2 vv = fmul2(cby[jj], dY);
                vv = ffma2(cbx[jj], dX, vv);
                vv = ffma2(ca[jj], dM, vv);
                const float2 mn = fmul2(pr, vv);
                const float2 xn = ffma2(ccx[jj], upX, upM);
                const float2 yn = ffma2(cg[jj], oY, oM);
                dM = oM; dX = oX; dY = oY;
                upM = mn; upX = xn;
                M[jj] = mn; X[jj] = xn; Y[jj] = yn;
            }
            bM = M[K - 1]; bX = X[K - 1]; bY = Y[K - 1];
            acc = fadd2(acc, ffma2(qi_last, bX, bM));
            --rem;
    };
    // Driver: lanes switch haplotype at different steps (lane t lags t columns), and a switch rewrites the whole
    // state, so a per-step "if (rem == 0)" is a scheduling barrier in every step.  Instead the warp asks how
    // many steps remain until ANY lane switches (one REDUX.MIN) and runs that many branch-free, four per loop
    // trip -- ptxas then overlaps the independent parts of consecutive columns; only the ~32 steps around each
    // haplotype boundary take the checked single-step path.
    int32_t s = 0;
    while (s < total) {
        const int32_t n = min((int32_t)__reduce_min_sync(0xffffffffu, rem), total - s);
        if (n >= HMM_FAST) {
            const int32_t nf = n - n % HMM_FAST;
#pragma unroll 1
            for (int32_t q = 0; q < nf; q += HMM_FAST) {
#pragma unroll
                for (int r_ = 0; r_ < HMM_FAST; ++r_) step();
            }
            s += nf;
            continue;
        }
        if (rem == 0) {
            // ---- this lane finished a haplotype: lane 31 owns the last row of both reads ----
            if (t == 31 && hidx >= h0) { sums_a[hidx - h0] = acc.x; sums_b[hidx - h0] = acc.y; }
            ++hidx;
            acc = zero2;
            if (hidx < h1) {
                const HapInfo hi = hapinfo[hidx];
                rem = hi.len;
                cp = codes + hi.off;
                inc = 1;
                init = hi.init;
                code_next = *cp;
            } else {
                rem = 0x7fffffff;    // drained: keep stepping on neutral input
                cp = zero_pad;
                inc = 0;
                code_next = 0u;
            }
#pragma unroll
            for (int jj = 0; jj < K; ++jj) {
                M[jj] = zero2; X[jj] = zero2;
                Y[jj] = make_float2((jj < n_pad_rows[0]) ? init : 0.f, (jj < n_pad_rows[1]) ? init : 0.f);
            }
            pdM = zero2; pdX = zero2;
            pdY = make_float2(top_boundary[0] ? init : 0.f, top_boundary[1] ? init : 0.f);
        }
        step();
        ++s;
    }
    // the last haplotype of lane 31 ends exactly at the last step
    if (t == 31 && hidx >= h0 && hidx < h1 && rem == 0) { sums_a[hidx - h0] = acc.x; sums_b[hidx - h0] = acc.y; }
}

// forward sums -> log10 likelihoods; sums FP32 cannot be trusted with are queued for the FP64 kernel
__global__ void __launch_bounds__(256)
hmm_finalize_kernel(HmmBatchView v, const int64_t *__restrict__ read_out_off,
                    const int32_t *__restrict__ order, int32_t n_items, const float *__restrict__ sums,
                    double *__restrict__ out, int2 *__restrict__ rescue, int32_t *__restrict__ rescue_count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    // order == nullptr: every read that fits the stream kernels (the rest is the striped kernel's)
    const int32_t r = order ? order[i] : i;
    if (!order && v.read_len[r] > 32 * HMM_MAX_K) return;
    const int32_t bt = v.read_batch[r];
    const int32_t h0 = (int32_t)v.batch_hap_start[bt], h1 = (int32_t)v.batch_hap_start[bt + 1];
    const int64_t o = read_out_off[r];
    const double lscale = log10((double)SCALE_F);
    for (int32_t h = h0; h < h1; ++h) {
        const float sfl = sums[o + (h - h0)];
        if (!(sfl >= RESCUE_BELOW) || !isfinite(sfl)) {
            rescue[atomicAdd(rescue_count, 1)] = make_int2(r, h);
            out[o + (h - h0)] = nan("");
        } else {
            out[o + (h - h0)] = log10((double)sfl) - lscale;
        }
    }
}

// ------------------------------------------------------------------------------------------
// generic striped kernel (FP64 exact-order path, long reads)
// ------------------------------------------------------------------------------------------
template <typename T> struct Num;
template <> struct Num<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float scale() { return SCALE_F; }
    static __device__ __forceinline__ bool bad(float s) { return !(s >= RESCUE_BELOW) || !isfinite(s); }
};
template <> struct Num<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double scale() { return DBL_MAX / 16; }
    static __device__ __forceinline__ bool bad(double) { return false; }
};

template <typename T, int K>
__device__ void hmm_striped_pair(const HmmBatchView &v, const double *lut, int32_t r, int64_t h,
                                 int gatk, T *scratch, double *out_slot, int2 *rescue,
                                 int32_t *rescue_count)
{
    const int t = threadIdx.x & 31;
    const int32_t R = v.read_len[r];
    const int32_t H = v.hap_len[h];
    const uint8_t *hap = v.buf + v.hap_off[h];
    const uint8_t *f_b = v.buf + v.read_field_off[5 * (int64_t)r + 0];
    const uint8_t *f_q = v.buf + v.read_field_off[5 * (int64_t)r + 1];
    const uint8_t *f_i = v.buf + v.read_field_off[5 * (int64_t)r + 2];
    const uint8_t *f_d = v.buf + v.read_field_off[5 * (int64_t)r + 3];
    const uint8_t *f_g = v.buf + v.read_field_off[5 * (int64_t)r + 4];
    constexpr int ROWS = 32 * K;
    const int n_stripes = (R + ROWS - 1) / ROWS;
    const int pad = n_stripes * ROWS - R;
    const T init = Num<T>::scale() / (T)H;
    T acc = (T)0;

    for (int st = 0; st < n_stripes; ++st) {
        T pm[K], px[K], cmm[K], cgm[K], cqi[K], cqd[K], cqg[K];
        int32_t rb[K];
        T M[K], X[K], Y[K];
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
            const int i = st * ROWS + t * K + jj - pad;
            if (i < 0) {
                pm[jj] = px[jj] = (T)0; cmm[jj] = cgm[jj] = cqi[jj] = cqd[jj] = (T)0; cqg[jj] = (T)1;
                rb[jj] = 0x100;
                Y[jj] = init;
            } else {
                const double Qr = lut[base_qual(f_q[i], gatk)], Qi = lut[f_i[i]], Qd = lut[f_d[i]], Qg = lut[f_g[i]];
                const int32_t base = f_b[i];
                pm[jj] = (T)(1 - Qr);
                px[jj] = (base == 'N') ? pm[jj] : (T)((gatk & 1) ? Qr / 3 : Qr);
                cmm[jj] = (T)(1 - (Qi + Qd));
                cgm[jj] = (T)(1 - Qg);
                cqi[jj] = (T)Qi; cqd[jj] = (T)Qd; cqg[jj] = (T)Qg;
                rb[jj] = base;
                Y[jj] = (T)0;
            }
            M[jj] = X[jj] = (T)0;
        }
        const bool first = (st == 0), last = (st == n_stripes - 1);
        const bool top_boundary = first && ((t * K - 1) < pad);
        T pdM = (T)0, pdX = (T)0, pdY = top_boundary ? init : (T)0;
        T bM = (T)0, bX = (T)0, bY = (T)0;
        T inM = (T)0, inX = (T)0, inY = (T)0;   // boundary row prefetched for lane 0
        const int S = H + 31;
        for (int s0 = 0; s0 < S; s0 += 32) {
            if (!first) {
                const int cc = s0 + t + 1;       // column this lane prefetches for lane 0
                if (cc <= H) {
                    inM = __ldcg(scratch + 3 * (int64_t)cc);
                    inX = __ldcg(scratch + 3 * (int64_t)cc + 1);
                    inY = __ldcg(scratch + 3 * (int64_t)cc + 2);
                }
            }
            const int send = min(32, S - s0);
            for (int u = 0; u < send; ++u) {
                const int s = s0 + u;
                const int c = s - t + 1;
                const bool active = (c >= 1 && c <= H);
                const int32_t hb = active ? (int32_t)__ldg(hap + c - 1) : 0x200;
                const bool hN = (hb == 'N');
                T upM = __shfl_up_sync(0xffffffffu, bM, 1);
                T upX = __shfl_up_sync(0xffffffffu, bX, 1);
                T upY = __shfl_up_sync(0xffffffffu, bY, 1);
                const T sM = __shfl_sync(0xffffffffu, inM, u);
                const T sX = __shfl_sync(0xffffffffu, inX, u);
                const T sY = __shfl_sync(0xffffffffu, inY, u);
                if (t == 0) {
                    if (first) { upM = (T)0; upX = (T)0; upY = init; }
                    else       { upM = sM; upX = sX; upY = sY; }
                }
                if (active) {
                    T dM = pdM, dX = pdX, dY = pdY;
                    pdM = upM; pdX = upX; pdY = upY;
#pragma unroll
                    for (int jj = 0; jj < K; ++jj) {
                        const T oM = M[jj], oX = X[jj], oY = Y[jj];
                        const T pr = (rb[jj] == hb || hN) ? pm[jj] : px[jj];
                        // reference association order, no FMA contraction
                        // (antidiagsPairHMM.c:184-195 / pairHMMmatrix.c:51-53)
                        const T mn = Num<T>::mul(pr, Num<T>::add(Num<T>::mul(cmm[jj], dM),
                                                                 Num<T>::mul(cgm[jj], Num<T>::add(dX, dY))));
                        const T xn = Num<T>::add(Num<T>::mul(upM, cqi[jj]), Num<T>::mul(upX, cqg[jj]));
                        const T yn = Num<T>::add(Num<T>::mul(oM, cqd[jj]), Num<T>::mul(oY, cqg[jj]));
                        dM = oM; dX = oX; dY = oY;
                        upM = mn; upX = xn;
                        M[jj] = mn; X[jj] = xn; Y[jj] = yn;
                    }
                    bM = M[K - 1]; bX = X[K - 1]; bY = Y[K - 1];
                    if (t == 31) {
                        if (last) {
                            acc = Num<T>::add(acc, Num<T>::add(bM, bX));
                        } else {
                            scratch[3 * (int64_t)c] = bM;
                            scratch[3 * (int64_t)c + 1] = bX;
                            scratch[3 * (int64_t)c + 2] = bY;
                        }
                    }
                }
            }
            __syncwarp();
        }
        __threadfence_block();
        __syncwarp();
    }
    if (t == 31) {
        if (Num<T>::bad(acc) && rescue != nullptr) {
            rescue[atomicAdd(rescue_count, 1)] = make_int2(r, (int)h);
            *out_slot = nan("");
        } else {
            *out_slot = log10((double)acc) - log10((double)Num<T>::scale());
        }
    }
}

// mode 0: items are explicit (read, haplotype) pairs; mode 1: items are reads, all haplotypes each
template <typename T, int K>
__global__ void __launch_bounds__(HMM_WARPS * 32)
hmm_striped_kernel(HmmBatchView v, const int64_t *__restrict__ read_out_off,
                   const int2 *__restrict__ pair_list, const int32_t *__restrict__ read_list,
                   const int32_t *__restrict__ n_items_ptr, int32_t n_items_host,
                   const double *__restrict__ lut_g, int gatk, T *__restrict__ scratch,
                   int64_t scratch_stride, double *__restrict__ out, int2 *__restrict__ rescue,
                   int32_t *__restrict__ rescue_count)
{
    __shared__ double lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = lut_g[i];
    __syncthreads();
    const int64_t warp = (int64_t)blockIdx.x * HMM_WARPS + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * HMM_WARPS;
    const int32_t n_items = n_items_ptr ? *n_items_ptr : n_items_host;
    T *my_scratch = scratch + warp * scratch_stride;
    for (int64_t it = warp; it < n_items; it += n_warps) {
        if (pair_list) {
            const int2 pr = pair_list[it];
            const int32_t bt = v.read_batch[pr.x];
            const int64_t o = read_out_off[pr.x] + (pr.y - v.batch_hap_start[bt]);
            hmm_striped_pair<T, K>(v, lut, pr.x, pr.y, gatk, my_scratch, out + o, rescue, rescue_count);
        } else {
            const int32_t r = read_list[it];
            const int32_t bt = v.read_batch[r];
            const int64_t h0 = v.batch_hap_start[bt], h1 = v.batch_hap_start[bt + 1];
            for (int64_t h = h0; h < h1; ++h)
                hmm_striped_pair<T, K>(v, lut, r, h, gatk, my_scratch, out + read_out_off[r] + (h - h0),
                                       rescue, rescue_count);
        }
    }
}

template <int K>
int launch_stream(const HmmBatchView &v, const int64_t *read_out_off, const int32_t *order,
                  int32_t count, const double *lut, int gatk, const uint8_t *codes, int64_t codes_bytes,
                  const HapInfo *info, float *sums, cudaStream_t st)
{
    if (count == 0) return AGX_OK;
    const int blocks = (count + HMM_WARPS - 1) / HMM_WARPS;
    hmm_stream_kernel<K><<<blocks, HMM_WARPS * 32, 0, st>>>(
        v, read_out_off, order + (int64_t)(K - 1) * v.n_reads, count, lut, gatk, codes, codes + codes_bytes,
        info, sums);
    count_launch();
    AGX_CUDA(cudaGetLastError());
    return AGX_OK;
}

template <int K>
int launch_duo(const HmmBatchView &v, const int64_t *read_out_off, const int2 *pairs, int32_t count,
               const double *lut, int gatk, const uint8_t *codes, int64_t codes_bytes, const HapInfo *info,
               float *sums, cudaStream_t st)
{
    if (count == 0) return AGX_OK;
    hmm_duo_kernel<K><<<count, 32, 0, st>>>(v, read_out_off, pairs + (int64_t)(K - 1) * v.n_reads, count, lut, gatk,
                                             codes, codes + codes_bytes, info, sums);
    count_launch();
    AGX_CUDA(cudaGetLastError());
    return AGX_OK;
}

}  // namespace

int hmm_workspace_reserve(HmmWorkspace &ws, int64_t n_reads, int64_t n_pairs, int64_t n_batches)
{
    if (n_batches > ws.cap_batches) {
        if (ws.batch_first) cudaFree(ws.batch_first);
        ws.batch_first = nullptr; ws.cap_batches = 0;
        AGX_CUDA(cudaMalloc(&ws.batch_first, (size_t)n_batches * 2 * sizeof(int32_t)));
        ws.cap_batches = n_batches;
    }
    if (!ws.counters) {
        AGX_CUDA(cudaMalloc(&ws.counters, HCNT_WORDS * sizeof(int32_t)));
        AGX_CUDA(cudaMallocHost(&ws.h_counters, HCNT_WORDS * sizeof(int32_t)));
        AGX_CUDA(cudaMalloc(&ws.d_lut, 256 * sizeof(double)));
        // Phred+33 -> probability with the HOST libm pow(), exactly the reference's expression
        // (antidiagsPairHMM.c:104-107); characters are `char` there, i.e. signed on x86.
        double lut[256];
        for (int c = 0; c < 256; ++c) lut[c] = pow(10.0, -((double)(signed char)c - 33.0) * 0.1);
        AGX_CUDA(cudaMemcpy(ws.d_lut, lut, sizeof lut, cudaMemcpyHostToDevice));
        for (int i = 0; i < 3; ++i) {
            AGX_CUDA(cudaStreamCreateWithFlags(&ws.aux[i], cudaStreamNonBlocking));
            AGX_CUDA(cudaEventCreateWithFlags(&ws.ev_join[i], cudaEventDisableTiming));
        }
        AGX_CUDA(cudaEventCreateWithFlags(&ws.ev_fork, cudaEventDisableTiming));
    }
    if (n_reads > ws.cap_reads) {
        if (ws.order) cudaFree(ws.order);
        ws.order = nullptr; ws.cap_reads = 0;
        AGX_CUDA(cudaMalloc(&ws.order, (size_t)n_reads * HMM_N_CLASSES * sizeof(int32_t)));
        if (ws.pairs) cudaFree(ws.pairs);
        ws.pairs = nullptr;
        AGX_CUDA(cudaMalloc(&ws.pairs, (size_t)n_reads * HMM_MAX_K * sizeof(int2)));
        ws.cap_reads = n_reads;
    }
    if (n_pairs > ws.cap_pairs) {
        if (ws.rescue) cudaFree(ws.rescue);
        ws.rescue = nullptr; ws.cap_pairs = 0;
        AGX_CUDA(cudaMalloc(&ws.rescue, (size_t)n_pairs * sizeof(int2)));
        ws.cap_pairs = n_pairs;
    }
    return AGX_OK;
}

void hmm_workspace_free(HmmWorkspace &ws)
{
    if (ws.order) cudaFree(ws.order);
    if (ws.pairs) cudaFree(ws.pairs);
    if (ws.batch_first) cudaFree(ws.batch_first);
    if (ws.counters) cudaFree(ws.counters);
    if (ws.h_counters) cudaFreeHost(ws.h_counters);
    if (ws.rescue) cudaFree(ws.rescue);
    if (ws.d_lut) cudaFree(ws.d_lut);
    if (ws.scratch) cudaFree(ws.scratch);
    if (ws.prep) cudaFree(ws.prep);
    ws.prof_stream.destroy(); ws.prof_fp64.destroy(); ws.prof_classify.destroy();
    for (int i = 0; i < 3; ++i) {
        if (ws.aux[i]) cudaStreamDestroy(ws.aux[i]);
        if (ws.ev_join[i]) cudaEventDestroy(ws.ev_join[i]);
    }
    if (ws.ev_fork) cudaEventDestroy(ws.ev_fork);
    ws = HmmWorkspace();
}

static int hmm_scratch_reserve(HmmWorkspace &ws, int64_t bytes)
{
    if (bytes > ws.cap_scratch) {
        if (ws.scratch) cudaFree(ws.scratch);
        ws.scratch = nullptr; ws.cap_scratch = 0;
        AGX_CUDA(cudaMalloc(&ws.scratch, (size_t)bytes));
        ws.cap_scratch = bytes;
    }
    return AGX_OK;
}

int hmm_run_device(HmmWorkspace &ws, const HmmBatchView &v, int64_t buf_bytes, const int64_t *d_read_out_off,
                   int64_t n_pairs, int gatk_mode, bool force_fp64, int do_rescue, double *d_out,
                   cudaStream_t st, cudaStream_t prep_st)
{
    if (v.n_reads == 0 || n_pairs == 0) return AGX_OK;
    ws.last_rescue = -1;
    if (v.n_reads > (int64_t)1 << 30 || v.n_haps > (int64_t)1 << 30)
        return fail(AGX_ERANGE, "pairhmm: more than 2^30 reads or haplotypes in one call");
    int rc = hmm_workspace_reserve(ws, v.n_reads, n_pairs, v.n_batches);
    if (rc != AGX_OK) return rc;

    // read pairing / row classes: on prep_st (high priority) when given, so that it is not queued behind the
    // stream kernels of another batch; the caller then guarantees ws is not in use by earlier work on st
    cudaStream_t cst = prep_st ? prep_st : st;
    const bool no_duo = getenv("AGX_PAIRHMM_NO_DUO") != nullptr;          // A/B switch: one read per warp only
    const int64_t nmax = v.n_reads > v.n_haps ? v.n_reads : v.n_haps;
    int2 *pairs = reinterpret_cast<int2 *>(ws.pairs);
    bool paired = !force_fp64 && !no_duo;
    ws.prof_classify.begin(cst);
    AGX_CUDA(cudaMemsetAsync(ws.counters, 0, HCNT_WORDS * sizeof(int32_t), cst));
    if (paired) {
        // reads of the same batch and row class are paired for hmm_duo_kernel
        int32_t *batch_first = ws.batch_first, *batch_end = ws.batch_first + v.n_batches;
        AGX_CUDA(cudaMemsetAsync(batch_first, 0xff, (size_t)v.n_batches * sizeof(int32_t), cst));
        AGX_CUDA(cudaMemsetAsync(batch_end, 0, (size_t)v.n_batches * sizeof(int32_t), cst));
        hmm_ranges_kernel<<<(int)((nmax + 255) / 256), 256, 0, cst>>>(v.read_batch, v.n_reads, v.hap_len, v.n_haps,
                                                                    batch_first, batch_end, ws.counters);
        const int64_t pair_blocks = std::min<int64_t>((v.n_batches + 3) / 4, 148 * 16);
        hmm_pair_kernel<<<(int)pair_blocks, 128, 0, cst>>>(v.read_len, v.n_reads, v.n_batches, batch_first, batch_end,
                                                        pairs, ws.order, ws.counters);
        count_launch(2);
        AGX_CUDA(cudaGetLastError());
        AGX_CUDA(cudaMemcpyAsync(ws.h_counters, ws.counters, HCNT_WORDS * sizeof(int32_t), cudaMemcpyDeviceToHost, cst));
        AGX_CUDA(cudaStreamSynchronize(cst));
        if (ws.h_counters[HCNT_UNSORTED]) {
            paired = false;               // reads of a batch are not consecutive: one read per warp
            AGX_CUDA(cudaMemsetAsync(ws.counters, 0, HCNT_WORDS * sizeof(int32_t), cst));
        }
    }
    if (!paired) {
        hmm_classify_kernel<<<(int)((nmax + 255) / 256), 256, 0, cst>>>(
            v.read_len, v.n_reads, v.hap_len, v.n_haps, ws.order, ws.counters, force_fp64 ? 1 : 0);
        count_launch();
        AGX_CUDA(cudaGetLastError());
        AGX_CUDA(cudaMemcpyAsync(ws.h_counters, ws.counters, HCNT_WORDS * sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, cst));
        AGX_CUDA(cudaStreamSynchronize(cst));
    }
    ws.prof_classify.end(cst);
    int32_t counts[HMM_N_CLASSES], pair_counts[HMM_MAX_K];
    for (int c = 0; c < HMM_N_CLASSES; ++c) counts[c] = ws.h_counters[c];
    for (int c = 0; c < HMM_MAX_K; ++c) pair_counts[c] = paired ? ws.h_counters[HCNT_PAIRS + c] : 0;
    const int32_t max_hap = ws.h_counters[HCNT_MAXHAP];
    const int gatk = gatk_mode & 3;     // bit 0: mismatch prior Qr/3, bit 1: base-quality floor 6 (GATK semantics)
    int2 *rescue = reinterpret_cast<int2 *>(ws.rescue);
    int32_t *rescue_count = ws.counters + HCNT_RESCUE;

    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t stride = 3 * ((int64_t)max_hap + 2);

    if (!force_fp64) {
        // haplotype symbol codes + per-haplotype records + FP32 forward sums live in one scratch block
        int64_t n_stream = 0;
        for (int c = 0; c < HMM_MAX_K; ++c) n_stream += counts[c] + 2 * (int64_t)pair_counts[c];
        uint8_t *codes = nullptr;
        const int64_t codes_bytes = ((buf_bytes + 255) / 256) * 256;   // a zeroed pad follows the codes
        HapInfo *info = nullptr;
        float *sums = nullptr;
        if (n_stream > 0) {
            const int64_t info_bytes = ((v.n_haps * (int64_t)sizeof(HapInfo) + 255) / 256) * 256;
            const int64_t sums_bytes = ((n_pairs * (int64_t)sizeof(float) + 255) / 256) * 256;
            const int64_t need = info_bytes + sums_bytes + codes_bytes + 256;
            if (need > ws.cap_prep) {
                if (ws.prep) cudaFree(ws.prep);
                ws.prep = nullptr; ws.cap_prep = 0;
                AGX_CUDA(cudaMalloc(&ws.prep, (size_t)need));
                ws.cap_prep = need;
            }
            info = reinterpret_cast<HapInfo *>(ws.prep);
            sums = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(ws.prep) + info_bytes);
            codes = reinterpret_cast<uint8_t *>(ws.prep) + info_bytes + sums_bytes;
            AGX_CUDA(cudaMemsetAsync(codes + codes_bytes, 0, 256, st));
            const int pblocks = (int)(v.n_haps < (int64_t)sms * 16 ? v.n_haps : (int64_t)sms * 16);
            hmm_prep_haps_kernel<<<pblocks, 128, 0, st>>>(v.buf, v.hap_off, v.hap_len, v.n_haps, codes, info);
            count_launch();
            AGX_CUDA(cudaGetLastError());
        }
        ws.prof_stream.begin(st);
        // longest rows first, launches spread round-robin over st and the auxiliary streams
        AGX_CUDA(cudaEventRecord(ws.ev_fork, st));
        int n_launched = 0;
        bool aux_used[3] = {false, false, false};
        auto next_stream = [&](cudaStream_t &out) -> int {
            const int slot = n_launched++ % 4;
            out = slot == 0 ? st : ws.aux[slot - 1];
            if (slot > 0 && !aux_used[slot - 1]) {
                AGX_CUDA(cudaStreamWaitEvent(out, ws.ev_fork, 0));
                aux_used[slot - 1] = true;
            }
            return AGX_OK;
        };
#define AGX_STREAM(KK)                                                                              \
    if (pair_counts[KK - 1] > 0) {                                                                  \
        cudaStream_t s_;                                                                            \
        if ((rc = next_stream(s_)) != AGX_OK) return rc;                                            \
        if ((rc = launch_duo<KK>(v, d_read_out_off, pairs, pair_counts[KK - 1], ws.d_lut, gatk,     \
                                 codes, codes_bytes, info, sums, s_)) != AGX_OK)                    \
            return rc;                                                                              \
    }                                                                                               \
    if (counts[KK - 1] > 0) {                                                                       \
        cudaStream_t s_;                                                                            \
        if ((rc = next_stream(s_)) != AGX_OK) return rc;                                            \
        if ((rc = launch_stream<KK>(v, d_read_out_off, ws.order, counts[KK - 1], ws.d_lut, gatk,    \
                                    codes, codes_bytes, info, sums, s_)) != AGX_OK)                 \
            return rc;                                                                              \
    }
        AGX_STREAM(8) AGX_STREAM(7) AGX_STREAM(6) AGX_STREAM(5)
        AGX_STREAM(4) AGX_STREAM(3) AGX_STREAM(2) AGX_STREAM(1)
#undef AGX_STREAM
        for (int i = 0; i < 3; ++i)
            if (aux_used[i]) {
                AGX_CUDA(cudaEventRecord(ws.ev_join[i], ws.aux[i]));
                AGX_CUDA(cudaStreamWaitEvent(st, ws.ev_join[i], 0));
            }
        ws.prof_stream.end(st);
        if (n_stream > 0) {
            hmm_finalize_kernel<<<(int)((v.n_reads + 255) / 256), 256, 0, st>>>(
                v, d_read_out_off, nullptr, (int32_t)v.n_reads, sums, d_out, rescue, rescue_count);
            count_launch();
            AGX_CUDA(cudaGetLastError());
        }
        if (counts[HMM_LONG] > 0) {
            // reads longer than 256 rows: FP32 striped kernel, one warp per read
            int64_t warps = counts[HMM_LONG];
            if (warps > (int64_t)sms * 16) warps = (int64_t)sms * 16;
            const int blocks = (int)((warps + HMM_WARPS - 1) / HMM_WARPS);
            rc = hmm_scratch_reserve(ws, stride * blocks * HMM_WARPS * (int64_t)sizeof(float));
            if (rc != AGX_OK) return rc;
            hmm_striped_kernel<float, 8><<<blocks, HMM_WARPS * 32, 0, st>>>(
                v, d_read_out_off, nullptr, ws.order + (int64_t)HMM_LONG * v.n_reads, nullptr,
                counts[HMM_LONG], ws.d_lut, gatk, reinterpret_cast<float *>(ws.scratch), stride, d_out,
                rescue, rescue_count);
            count_launch();
            AGX_CUDA(cudaGetLastError());
        }
        if (!do_rescue) return AGX_OK;
        int32_t n_rescue = 0;
        if (do_rescue == 1) {
            AGX_CUDA(cudaMemcpyAsync(ws.h_counters + HCNT_RESCUE, rescue_count, sizeof(int32_t),
                                     cudaMemcpyDeviceToHost, st));
            AGX_CUDA(cudaStreamSynchronize(st));
            n_rescue = ws.h_counters[HCNT_RESCUE];
            ws.last_rescue = n_rescue;
            if (n_rescue == 0) return AGX_OK;
        } else {
            n_rescue = (int32_t)std::min<int64_t>(n_pairs, (int64_t)sms * 4);   // grid only; the count stays on the device
        }
        int64_t warps = n_rescue;
        if (warps > (int64_t)sms * 16) warps = (int64_t)sms * 16;
        const int blocks = (int)((warps + HMM_WARPS - 1) / HMM_WARPS);
        rc = hmm_scratch_reserve(ws, stride * blocks * HMM_WARPS * (int64_t)sizeof(double));
        if (rc != AGX_OK) return rc;
        ws.prof_fp64.begin(st);
        hmm_striped_kernel<double, 4><<<blocks, HMM_WARPS * 32, 0, st>>>(
            v, d_read_out_off, rescue, nullptr, do_rescue == 2 ? rescue_count : nullptr, n_rescue, ws.d_lut, gatk,
            reinterpret_cast<double *>(ws.scratch), stride, d_out, nullptr, nullptr);
        ws.prof_fp64.end(st);
        count_launch();
        AGX_CUDA(cudaGetLastError());
        return AGX_OK;
    }

    // force_fp64: every read goes through the exact-order FP64 kernel (classified "long" above)
    int64_t warps = counts[HMM_LONG];
    if (warps > (int64_t)sms * 16) warps = (int64_t)sms * 16;
    const int blocks = (int)((warps + HMM_WARPS - 1) / HMM_WARPS);
    if (blocks == 0) return AGX_OK;
    rc = hmm_scratch_reserve(ws, stride * blocks * HMM_WARPS * (int64_t)sizeof(double));
    if (rc != AGX_OK) return rc;
    hmm_striped_kernel<double, 4><<<blocks, HMM_WARPS * 32, 0, st>>>(
        v, d_read_out_off, nullptr, ws.order + (int64_t)HMM_LONG * v.n_reads, nullptr, counts[HMM_LONG],
        ws.d_lut, gatk, reinterpret_cast<double *>(ws.scratch), stride, d_out, nullptr, nullptr);
    count_launch();
    AGX_CUDA(cudaGetLastError());
    return AGX_OK;
}

}  // namespace agx
