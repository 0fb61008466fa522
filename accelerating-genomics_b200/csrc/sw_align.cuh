// sw_align.cuh -- alignment END CELL, START CELL and CIGAR on top of the Smith-Waterman kernels (included by
// sw_kernels.cu inside namespace agx { namespace { ).
//
// the reference leaves traceback out (its README), so only the END CELL has a
// reference meaning -- the cell its running maximum comes from (antidiagonalSmithWaterman.c:335, strict `>`,
// cells visited by anti-diagonals :270-347 with ix, the shorter line / line 1 on ties :229-244, ascending).  START
// and CIGAR follow the rule stated in oracle/sw_align.c (diagonal first, then the shortest gap, line 1's symbols
// before line 2's), which needs the H matrix alone.
//
// Data flow of one call (all on the device):
//   sw_align_classify_kernel   orientation, length class, most CIGAR runs a pair can have
//   sw_duo_kernel<G,K,MODE>    MODE 1: score + end cell from per-row 32-bit keys; MODE 2: also the LOW BYTE of
//   sw_wave_kernel<MODE>       every H (1 byte per cell: neighbouring H differ by less than 128, so the walk
//                              rebuilds exact values from byte differences)
//   sw_walk_kernel             one thread per pair walks end -> start through the byte matrix, emits runs
//   exclusive scan + sw_cigar_gather_kernel   runs reversed into one packed array

constexpr int CNT_A_ROWS = SW_N_CLASSES + 4;                 // [SW_N_DUO_CLASSES] most rows of a duo class
constexpr int CNT_A_BAD = CNT_A_ROWS + SW_N_DUO_CLASSES;     // pairs outside the supported range
constexpr int CNT_A_WORDS = CNT_A_BAD + 2;

constexpr int32_t ALIGN_DUO_MAX_ROWS = 2047;     // 32-bit key: 12 bits of anti-diagonal
constexpr int32_t ALIGN_WAVE_MAX_COLS = 16000;   // 64-bit key: 14 bits of column

__global__ void __launch_bounds__(256)
sw_align_classify_kernel(const uint8_t *__restrict__ seqs, const int64_t *__restrict__ off,
                         const int32_t *__restrict__ len, int64_t n_pairs, int32_t s16_max_short, int64_t long_cells,
                         int32_t match, int mode, int32_t *__restrict__ order, int32_t *__restrict__ counters,
                         int32_t *__restrict__ scores, int32_t *__restrict__ ends, int32_t *__restrict__ cap32,
                         SwWalkRec *__restrict__ wk)
{
    __shared__ int32_t s_cnt[SW_N_CLASSES];
    __shared__ int32_t s_base[SW_N_CLASSES];
    __shared__ int32_t s_rows[SW_N_DUO_CLASSES];
    __shared__ int32_t s_maxlen, s_bad;
    if (threadIdx.x < SW_N_CLASSES) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x < SW_N_DUO_CLASSES) s_rows[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_maxlen = 0; s_bad = 0; }
    __syncthreads();

    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int cls = -1, rank = 0;
    if (p < n_pairs) {
        const int32_t rx = len[2 * p], ry = len[2 * p + 1];
        int32_t lx = rx, ly = ry;
        const bool a_is_x = !(rx > ry);                      // antidiagonalSmithWaterman.c:229
        const bool nx = strip_newline(seqs + off[2 * p], lx);
        const bool ny = strip_newline(seqs + off[2 * p + 1], ly);
        const int32_t la = a_is_x ? lx : ly, lb = a_is_x ? ly : lx;      // columns / rows, newline symbols off
        const int32_t ra = a_is_x ? rx : ry, rb = a_is_x ? ry : rx;
        if (mode == 2) cap32[p] = rx + ry + 2;
        if (lx == 0 || ly == 0) {
            const int32_t s = (nx && ny) ? match : 0;
            scores[p] = s;
            ends[2 * p] = s > 0 ? lx : -1;
            ends[2 * p + 1] = s > 0 ? ly : -1;
            if (mode == 2) {
                SwWalkRec w = {};
                w.flags = (uint8_t)(SW_WK_TRIVIAL | (s > 0 ? 0 : SW_WK_NONE));
                wk[p] = w;
            }
        } else {
            const int64_t stripes = (ra + 255) / 256;
            if (la <= s16_max_short && lb <= ALIGN_DUO_MAX_ROWS) {
#pragma unroll
                for (int c = SW_N_DUO_CLASSES - 1; c >= 0; --c)
                    if (la <= duo_cap(c)) cls = c;
                atomicMax(&s_rows[cls], lb);
            } else if (ra <= ALIGN_WAVE_MAX_COLS && (int64_t)ra + rb < (int64_t)WAVE_DMAX &&
                       stripes * rb < ((int64_t)1 << 30) && !((int64_t)rx * ry >= long_cells && min(lx, ly) > DUO_MAX_CAP)) {
                cls = GENERIC;
            } else {
                atomicAdd(&s_bad, 1);
            }
            if (cls >= 0) {
                rank = atomicAdd(&s_cnt[cls], 1);
                atomicMax(&s_maxlen, max(rx, ry));
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < SW_N_CLASSES && s_cnt[threadIdx.x] > 0)
        s_base[threadIdx.x] = atomicAdd(&counters[threadIdx.x], s_cnt[threadIdx.x]);
    if (threadIdx.x < SW_N_DUO_CLASSES && s_rows[threadIdx.x] > 0) atomicMax(&counters[CNT_A_ROWS + threadIdx.x], s_rows[threadIdx.x]);
    if (threadIdx.x == 0 && s_maxlen > 0) atomicMax(&counters[CNT_MAXLEN], s_maxlen);
    if (threadIdx.x == 0 && s_bad > 0) atomicAdd(&counters[CNT_A_BAD], s_bad);
    __syncthreads();
    if (cls >= 0) order[(int64_t)cls * n_pairs + s_base[cls] + rank] = (int32_t)p;
}

// 256-byte units of the H-byte matrix of every listed wavefront pair: stripes x rows (raw lengths)
__global__ void __launch_bounds__(256)
sw_wave_tb_units_kernel(const int32_t *__restrict__ len, const int32_t *__restrict__ list, const int32_t *__restrict__ list_count,
                        int32_t *__restrict__ units, int64_t cap)
{
    const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= cap) return;
    int32_t u = 0;
    if (it < *list_count) {
        const int32_t p = list[it];
        const int32_t rx = len[2 * (int64_t)p], ry = len[2 * (int64_t)p + 1];
        const int32_t ra = rx > ry ? ry : rx, rb = rx > ry ? rx : ry;
        u = ((ra + 255) / 256) * rb;
    }
    units[it] = u;
}
__global__ void __launch_bounds__(256) units_to_bytes_kernel(int64_t *__restrict__ off, int64_t n, int64_t base)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) off[i] = off[i] * 256 + base;
}

// ---- the traceback walk ---------------------------------------------------------------------------------------
struct WalkLayout {
    const uint8_t *tb;      // the pair's matrix
    int32_t rstride, K, TB, LS, half;
    bool duo;
    __device__ __forceinline__ uint32_t at(int32_t r, int32_t c) const
    {
        if (duo) {
            // lane t of the sub-warp holds row r at step r + t; batches of TB steps, a lane block of LS words per
            // batch, word (q, step in batch) at q * TB + step (see duo_tb_steps)
            const int32_t t = c / K, jj = c - t * K, s = r + t;
            return __ldg(tb + ((int64_t)(s / TB) * 32 + t) * (LS * 4) + ((jj >> 1) * TB + (s & (TB - 1))) * 4 + (jj & 1) + 2 * half);
        }
        return __ldg(tb + ((int64_t)(c >> 8) * rstride + r) * 256 + (c & 255));
    }
};

// Pairs ordered by the size of their score (12 bins of floor(log2)): the length of a walk goes with the score, and a warp
// whose 32 walks end together keeps 32 loads in flight instead of the 14 it averaged on a half related / half
// unrelated batch.  A counting sort in two passes; the order inside a bin does not matter.
constexpr int WALK_BINS = 12;
__device__ __forceinline__ int walk_bin(int32_t score) { return score <= 0 ? 0 : min(WALK_BINS - 1, 32 - __clz(score)); }
__global__ void __launch_bounds__(256)
sw_walk_count_kernel(const int32_t *__restrict__ scores, int64_t n_pairs, int32_t *__restrict__ bins)
{
    __shared__ int32_t s_cnt[WALK_BINS];
    if (threadIdx.x < WALK_BINS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_pairs) atomicAdd(&s_cnt[walk_bin(scores[p])], 1);
    __syncthreads();
    if (threadIdx.x < WALK_BINS && s_cnt[threadIdx.x] > 0) atomicAdd(&bins[threadIdx.x], s_cnt[threadIdx.x]);
}
__global__ void __launch_bounds__(256)
sw_walk_order_kernel(const int32_t *__restrict__ scores, int64_t n_pairs, int32_t *__restrict__ bins, int32_t *__restrict__ perm)
{
    // bins[0 .. WALK_BINS) = counts, bins[WALK_BINS .. 2 WALK_BINS) = cursors (zero on entry); largest scores first
    __shared__ int32_t s_cnt[WALK_BINS], s_base[WALK_BINS];
    if (threadIdx.x < WALK_BINS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int b = -1, rank = 0;
    if (p < n_pairs) { b = walk_bin(scores[p]); rank = atomicAdd(&s_cnt[b], 1); }
    __syncthreads();
    if (threadIdx.x < WALK_BINS) {
        int32_t first = 0;
        for (int k = WALK_BINS - 1; k > (int)threadIdx.x; --k) first += bins[k];
        s_base[threadIdx.x] = first + (s_cnt[threadIdx.x] > 0 ? atomicAdd(&bins[WALK_BINS + threadIdx.x], s_cnt[threadIdx.x]) : 0);
    }
    __syncthreads();
    if (b >= 0) perm[s_base[b] + rank] = (int32_t)p;
}

__global__ void __launch_bounds__(128)
sw_walk_kernel(const uint8_t *__restrict__ seqs, const int64_t *__restrict__ off, const int32_t *__restrict__ len,
               int64_t n_pairs, SwScoring sc, const SwWalkRec *__restrict__ wk, const uint8_t *__restrict__ tb,
               const uint8_t *__restrict__ tb_gen, const int32_t *__restrict__ scores, const int32_t *__restrict__ ends, int32_t *__restrict__ coords,
               uint32_t *__restrict__ tmp_ops, const int64_t *__restrict__ tmp_off, int32_t *__restrict__ nops,
               int32_t *__restrict__ bad, const int32_t *__restrict__ perm)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pairs) return;
    const int64_t p = perm[idx];
    const SwWalkRec w = wk[p];
    const int32_t score = scores[p];
    uint32_t *out = tmp_ops + tmp_off[p];
    int32_t *co = coords + 4 * p;
    const int32_t ea = ends[2 * p], eb = ends[2 * p + 1];
    if (w.flags & SW_WK_NONE) {
        co[0] = co[1] = co[2] = co[3] = -1;
        nops[p] = 0;
        return;
    }
    if (w.flags & SW_WK_TRIVIAL) {
        co[0] = co[1] = ea;
        co[2] = co[3] = eb;
        out[0] = (1u << 4) | 0u;
        nops[p] = 1;
        return;
    }
    const bool a_is_x = (w.flags & SW_WK_A_IS_X) != 0;
    const uint8_t *cs = seqs + off[2 * p + (a_is_x ? 0 : 1)];      // column sequence
    const uint8_t *rs = seqs + off[2 * p + (a_is_x ? 1 : 0)];      // row sequence
    WalkLayout L;
    L.duo = w.cls >= 0;
    L.tb = (L.duo ? tb : tb_gen) + w.tb_off;
    L.rstride = w.rstride;
    L.K = L.duo ? duo_class(w.cls).k : 1;
    L.TB = duo_tb_steps(L.K);
    L.LS = duo_tb_block_words(L.K);
    L.half = w.half;
    const int32_t goe_open = sc.gap_open, ext = sc.gap_extend;
    // runs, end -> start, merged as they come
    int32_t n_runs = 0, cur_op = -1, cur_len = 0;
    auto emit = [&](int32_t op, int32_t n) {
        if (op == cur_op) { cur_len += n; return; }
        if (cur_op >= 0) out[n_runs++] = ((uint32_t)cur_len << 4) | (uint32_t)cur_op;
        cur_op = op;
        cur_len = n;
    };
    // a step to the left consumes a symbol of the column sequence, a step up one of the row sequence;
    // I = symbols of line 1 only, D = symbols of line 2 only
    const int32_t op_left = a_is_x ? 1 : 2, op_up = a_is_x ? 2 : 1;
    int32_t r = w.r_end, c = w.c_end, h = score;
    if (w.flags & SW_WK_NL_END) { emit(0, 1); h -= sc.match; }
    int32_t i = r - w.row_off, j = c - w.col_off;                   // symbol indices: rs[i], cs[j]
    uint32_t bc = L.at(r, c);
    bool broken = false;
    constexpr int WD = 4;     // cells of the diagonal requested together: a walk is one dependent load after the other
    while (h > 0) {
        // the diagonal run first: (r-1, c-1) .. (r-WD, c-WD) and their symbols are loaded side by side, then checked
        // in order -- the same decisions as one cell at a time, a quarter of the round trips
        uint32_t bdv[WD], csv[WD], rsv[WD];
#pragma unroll
        for (int d = 0; d < WD; ++d) {
            const bool in = i - d > 0 && j - d > 0;
            bdv[d] = in ? L.at(r - d - 1, c - d - 1) : 0u;
            const bool ins = i - d >= 0 && j - d >= 0;
            csv[d] = ins ? cs[j - d] : 0u;
            rsv[d] = ins ? rs[i - d] : 1u;
        }
        bool off_diagonal = false;
#pragma unroll
        for (int d = 0; d < WD; ++d) {
            if (h <= 0 || off_diagonal) break;
            const int32_t hd = (i > 0 && j > 0) ? h + (int32_t)(int8_t)(uint8_t)(bdv[d] - bc) : 0;
            if (h == hd + (csv[d] == rsv[d] ? sc.match : sc.mismatch)) {
                emit(0, 1);
                --r; --c; --i; --j;
                h = hd;
                bc = bdv[d];
            } else {
                off_diagonal = true;
            }
        }
        if (!off_diagonal) continue;
        int32_t hl = h, hu = h;            // exact H while walking left / up
        uint32_t bl = bc, bu = bc;
        bool found = false;
        for (int32_t k = 1; !found; ++k) {
            const bool can_l = j - k >= 0, can_u = i - k >= 0;
            if (!can_l && !can_u) break;
            const int32_t want = h - goe_open - k * ext;          // H of the cell the gap leaves from
#pragma unroll
            for (int pass = 0; pass < 2 && !found; ++pass) {
                const bool left = (pass == 0) == a_is_x;          // line 1's symbols first
                if (left) {
                    if (!can_l) continue;
                    const uint32_t bn = L.at(r, c - k);
                    hl += (int32_t)(int8_t)(uint8_t)(bn - bl);
                    bl = bn;
                    if (hl == want) { emit(op_left, k); c -= k; j -= k; h = hl; bc = bl; found = true; }
                } else {
                    if (!can_u) continue;
                    const uint32_t bn = L.at(r - k, c);
                    hu += (int32_t)(int8_t)(uint8_t)(bn - bu);
                    bu = bn;
                    if (hu == want) { emit(op_up, k); r -= k; i -= k; h = hu; bc = bu; found = true; }
                }
            }
        }
        if (!found) { broken = true; break; }
    }
    if (cur_op >= 0) out[n_runs++] = ((uint32_t)cur_len << 4) | (uint32_t)cur_op;
    if (broken) { atomicAdd(bad, 1); n_runs = 0; }
    nops[p] = n_runs;
    // first aligned symbols: the cell after the one the walk stopped in
    const int32_t sa = j + 1, sb = i + 1;
    co[0] = a_is_x ? sa : sb;
    co[1] = ea;
    co[2] = a_is_x ? sb : sa;
    co[3] = eb;
}

__global__ void __launch_bounds__(256)
sw_cigar_gather_kernel(const uint32_t *__restrict__ tmp_ops, const int64_t *__restrict__ tmp_off, const int32_t *__restrict__ nops,
                       const int64_t *__restrict__ cig_off, int64_t n_pairs, uint32_t *__restrict__ cigar)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const int32_t n = nops[p];
    const uint32_t *src = tmp_ops + tmp_off[p];
    uint32_t *dst = cigar + cig_off[p];
    for (int32_t k = 0; k < n; ++k) dst[k] = src[n - 1 - k];
}

// bytes of one warp's step-major matrix in class c when the class's longest row sequence has `rows` symbols
__host__ __device__ constexpr int64_t duo_tb_warp_bytes(int c, int32_t rows)
{
    const int tb = duo_tb_steps(duo_class(c).k);
    return (int64_t)((rows + duo_class(c).g - 1 + tb - 1) / tb) * duo_tb_batch_bytes(duo_class(c).k);
}
// ... of the whole class: the grid is whole CTAs of DUO_THREADS / 32 warps
inline int64_t duo_tb_class_bytes(int c, int32_t count, int32_t rows)
{
    const int subs = DUO_THREADS / duo_class(c).g;
    const int64_t duos = (count + 1) / 2, blocks = (duos + subs - 1) / subs;
    return (blocks * (DUO_THREADS / 32) * duo_tb_warp_bytes(c, rows) + 255) / 256 * 256;
}

template <int C>
int launch_duo_align(const uint8_t *d_seqs, const int64_t *d_off, const int32_t *d_len, const int32_t *order,
                     int64_t n_pairs, const int32_t *counts, const int32_t *rows, const int64_t *tb_base, int mode,
                     const DuoConst &kc, int32_t *d_scores, int32_t *counters, DuoAlignOut ao, cudaStream_t st)
{
    if constexpr (C < SW_N_DUO_CLASSES) {
        constexpr int G = duo_class(C).g, K = duo_class(C).k;
        constexpr int SUBS = DUO_THREADS / G;
        const int duos = (counts[C] + 1) / 2;
        const int blocks = (duos + SUBS - 1) / SUBS;
        if (blocks > 0) {
            DuoAlignOut a = ao;
            a.cls = (int16_t)C;
            a.tb_warp_bytes = duo_tb_warp_bytes(C, rows[C]);
            a.tb = ao.tb ? ao.tb + tb_base[C] : nullptr;
            a.tb_class_off = tb_base[C];
            int32_t *glist = const_cast<int32_t *>(order) + (int64_t)GENERIC * n_pairs;
            if (mode == 2) {
                // staging buffers in dynamic shared memory (with the 8 KB row ring they pass the 48 KB static limit)
                AGX_CUDA(cudaFuncSetAttribute(sw_duo_kernel<G, K, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, duo_tb_smem_bytes(K)));
                sw_duo_kernel<G, K, 2><<<blocks, DUO_THREADS, duo_tb_smem_bytes(K), st>>>(d_seqs, d_off, d_len, order + (int64_t)C * n_pairs,
                                                                     counts[C], kc, d_scores, glist, counters + GENERIC, a);
            }
            else
                sw_duo_kernel<G, K, 1><<<blocks, DUO_THREADS, 0, st>>>(d_seqs, d_off, d_len, order + (int64_t)C * n_pairs,
                                                                     counts[C], kc, d_scores, glist, counters + GENERIC, a);
            count_launch();
            AGX_CUDA(cudaGetLastError());
        }
        return launch_duo_align<C + 1>(d_seqs, d_off, d_len, order, n_pairs, counts, rows, tb_base, mode, kc, d_scores,
                                       counters, ao, st);
    } else {
        return AGX_OK;
    }
}
