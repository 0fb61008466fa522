// sw_long.cu -- one very long Smith-Waterman alignment spread over every SM of one or several B200s.
//
// Same recurrence as sw_kernels.cu (antidiagonalSmithWaterman.c:290-335), raw-byte comparison, s32
// DPX.  The DP matrix is cut into column stripes of 32*K columns.  A warp owns one stripe at a time
// and sweeps all rows of it systolically (lane t holds K columns in registers and is one row behind
// lane t-1, exactly like sw_wave_kernel); the stripes themselves form a second, coarser wavefront:
// stripe j may process row block i as soon as stripe j-1 has published the right boundary column
// (H+goe and E per row) of that block.
//
//   - ONE boundary array of 2*rows int32 per GPU is shared by all stripes and updated in place: the
//     region of row block i always holds the boundary of the last stripe that passed it, and stripes
//     pass a block strictly in order.
//   - progress[j] (one int per stripe) counts the row blocks whose boundary stripe j-1 has published for
//     stripe j; a warp spins on it with __nanosleep back-off.  All warps are co-resident (cooperative
//     launch), so the wait always ends.
//   - Multi-GPU: GPU g owns a contiguous range of columns.  The last stripe of GPU g writes its
//     boundary column straight into GPU g+1's boundary array and bumps GPU g+1's progress[0] with
//     peer stores over NVLink (cudaDeviceEnablePeerAccess); no collective is involved, the final
//     score is the max of the per-GPU maxima.
#include <algorithm>
#include <cstdlib>
#include <thread>
#include <vector>

#include "common.cuh"

namespace agx {

namespace {

constexpr int LONG_WARPS = 4;     // warps per CTA
constexpr int LONG_RING = 64;
constexpr int LONG_RB_DEFAULT = 32;   // rows per published block (multiple of 32)

struct LongArgs {
    const uint8_t *a;         // columns owned by this GPU
    int32_t la;
    const uint8_t *b;         // rows
    int32_t lb;
    int32_t *bnd;             // [2 * lb] local boundary array (in place)
    int32_t *progress;        // [n_stripes + 1]; progress[j] gates stripe j
    int32_t *next_bnd;        // boundary array of the next GPU (peer) or nullptr
    int32_t *next_progress;   // &progress[0] of the next GPU (peer) or nullptr
    int32_t *best;            // running maximum (atomicMax)
    int32_t first_gpu;        // stripe 0 of this GPU is the true left edge of the matrix
    int32_t rb;               // rows per published block (multiple of 32)
    SwScoring sc;
};

// acquire loads of a progress counter: .gpu for a counter written on this GPU, .sys for the one a peer
// GPU bumps over NVLink
__device__ __forceinline__ int32_t ld_acquire(const int32_t *p, bool sys)
{
    int32_t v;
    if (sys) asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else     asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int K, bool SHORT>
__global__ void __launch_bounds__(LONG_WARPS * 32)
sw_long_kernel(LongArgs g)
{
    constexpr int W = 32 * K;
    __shared__ int32_t r_byte[LONG_WARPS][LONG_RING];
    __shared__ int32_t r_g[LONG_WARPS][LONG_RING];
    __shared__ int32_t r_e[LONG_WARPS][LONG_RING];

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * LONG_WARPS + wib;
    const int n_warps = gridDim.x * LONG_WARPS;
    const int32_t goe = g.sc.gap_open + g.sc.gap_extend;
    const int32_t ext = g.sc.gap_extend;
    const int32_t sub_match = g.sc.match - goe, sub_mis = g.sc.mismatch - goe;
    const int32_t lb = g.lb;
    const int n_stripes = (g.la + W - 1) / W;
    const int LONG_RB = g.rb;
    const int n_blocks = (lb + LONG_RB - 1) / LONG_RB;
    int32_t bestg = goe;          // running max of H + goe

    for (int st = warp; st < n_stripes; st += n_warps) {
        const int c0 = st * W + lane * K;
        int32_t acol[K], Gp[K], F[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            acol[j] = (c0 + j < g.la) ? (int32_t)g.a[c0 + j] : 0x100;
            Gp[j] = goe;
            F[j] = goe;
        }
        int32_t g_out = goe, e_out = goe, g_in_prev = goe;
        const bool left_edge = (st == 0) && g.first_gpu;
        const bool last = (st == n_stripes - 1);
        int32_t *out_bnd = last ? g.next_bnd : g.bnd;                    // nullptr: nothing to publish
        int32_t *out_prog = last ? g.next_progress : (g.progress + st + 1);
        const bool remote = last && g.next_bnd != nullptr;
        const bool gate_is_remote = (st == 0) && !g.first_gpu;
        const int S = lb + 31;
        int published = 0;
        for (int s0 = 0; s0 < S; s0 += 32) {
            // publish the row blocks whose boundary writes are complete (rows < s0 - 31)
            if (out_bnd != nullptr) {
                const int done = (s0 >= 32) ? min(n_blocks, (s0 - 32) / LONG_RB) : 0;
                if (done > published) {
                    __syncwarp();
                    if (lane == 31) {
                        if (remote) __threadfence_system(); else __threadfence();
                        *((volatile int32_t *)out_prog) = done;
                    }
                    published = done;
                }
            }
            // wait until the left neighbour has published the block these 32 rows belong to
            if (!left_edge && s0 < lb) {
                const int need = min(n_blocks, s0 / LONG_RB + 1);
                if (lane == 0) {
                    unsigned ns = 32;
                    while (ld_acquire(g.progress + st, gate_is_remote) < need) {
                        __nanosleep(ns);
                        if (ns < 1024) ns *= 2;
                    }
                }
                __syncwarp();
            }
            {
                const int r = s0 + lane;
                int32_t bb = 0x200, gi = goe, ei = goe;
                if (r < lb) {
                    bb = g.b[r];
                    if (!left_edge) { gi = __ldcg(g.bnd + 2 * (int64_t)r); ei = __ldcg(g.bnd + 2 * (int64_t)r + 1); }
                }
                r_byte[wib][r & (LONG_RING - 1)] = bb;
                r_g[wib][r & (LONG_RING - 1)] = gi;
                r_e[wib][r & (LONG_RING - 1)] = ei;
            }
            __syncwarp();
            const int send = min(32, S - s0);
#pragma unroll 1
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
                        const int32_t tg = __vimax_s32_relu(F[j], d) + goe;      // T[j] + goe
                        e = __viaddmax_s32(e, ext, tg_prev);                       // E[i][j]
                        gdiag = Gp[j];
                        gleft = max(e + goe, tg);                                  // H[i][j] + goe
                        Gp[j] = gleft;
                        tg_prev = tg;
                        bestg = max(bestg, gleft);
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
                        bestg = max(bestg, gleft);
                    }
                }
                g_out = gleft;
                e_out = e;
                const int r = s - 31;
                if (out_bnd != nullptr && lane == 31 && r >= 0 && r < lb) {
                    out_bnd[2 * (int64_t)r] = g_out;
                    out_bnd[2 * (int64_t)r + 1] = e_out;
                }
            }
        }
        // every row of this stripe is written: publish the last blocks
        if (out_bnd != nullptr) {
            __syncwarp();
            if (lane == 31) {
                if (remote) __threadfence_system(); else __threadfence();
                *((volatile int32_t *)out_prog) = n_blocks;
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) bestg = max(bestg, __shfl_xor_sync(0xffffffffu, bestg, m));
    const int32_t best = bestg - goe;
    if (lane == 0 && best > 0) atomicMax(g.best, best);
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
    return (e && atoi(e) > 0) ? atoi(e) : dflt;
}

// Stripe width.  The stripe wavefront costs (total stripes x hop + rows) steps, so stripes should be as
// wide as the register file allows while every SM still gets a few of them.
int pick_k(int64_t cols_per_gpu, int sms)
{
    const int forced = env_int("AGX_LONG_K", 0);
    if (forced == 2 || forced == 4 || forced == 8 || forced == 16 || forced == 32) return forced;
    for (int k : {32, 16, 8, 4}) {
        const int64_t stripes = (cols_per_gpu + 32 * k - 1) / (32 * k);
        if (stripes >= (int64_t)sms * 2) return k;       // at least two stripes per SM
    }
    return 2;
}

int pick_rb() { return (env_int("AGX_LONG_RB", LONG_RB_DEFAULT) + 31) / 32 * 32; }

int long_dispatch(int k, const LongArgs &args, cudaStream_t st)
{
    const int n = (args.la + 32 * k - 1) / (32 * k);
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // few warps per SM: latency-bound, take the short E chain; many: issue-bound, take the lean one
    int short_chain = (n < sms * 4) ? 1 : 0;
    if (const char *e = getenv("AGX_LONG_CHAIN")) short_chain = atoi(e) != 0;
#define AGX_LONG_CASE(KK)                                                                     \
    case KK: return short_chain ? long_launch<KK, true>(args, n, st) : long_launch<KK, false>(args, n, st);
    switch (k) {
        AGX_LONG_CASE(32)
        AGX_LONG_CASE(16)
        AGX_LONG_CASE(8)
        AGX_LONG_CASE(4)
    default: return short_chain ? long_launch<2, true>(args, n, st) : long_launch<2, false>(args, n, st);
    }
#undef AGX_LONG_CASE
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
    const int k = pick_k(la, sms);
    const int64_t n_stripes = (la + 32 * k - 1) / (32 * k);
    const int64_t need = 2 * lb + n_stripes + 2;
    if (need > ws.cap) {
        if (ws.buf) cudaFree(ws.buf);
        ws.buf = nullptr; ws.cap = 0;
        AGX_CUDA(cudaMalloc(&ws.buf, (size_t)need * sizeof(int32_t)));
        ws.cap = need;
    }
    LongArgs args;
    args.a = d_a; args.la = (int32_t)la; args.b = d_b; args.lb = (int32_t)lb;
    args.bnd = ws.buf;
    args.progress = ws.buf + 2 * lb;
    args.next_bnd = nullptr; args.next_progress = nullptr;
    args.best = d_best;
    args.first_gpu = 1;
    args.rb = pick_rb();
    args.sc = sc;
    AGX_CUDA(cudaMemsetAsync(args.progress, 0, (size_t)(n_stripes + 2) * sizeof(int32_t), st));
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
    // interior cuts sit on multiples of the widest stripe (32 lanes x 32 columns): only the very last
    // stripe of the matrix may carry padding columns, whose boundary nobody consumes
    for (int gidx = 0; gidx <= n_dev; ++gidx) c_lo[gidx] = (gidx == n_dev) ? la : (la * gidx / n_dev) / 1024 * 1024;
    // allocate + upload on every GPU, clear the flags, then make sure ALL GPUs are clear before any launch
    for (int gidx = 0; gidx < n_dev; ++gidx) {
        AGX_CUDA(cudaSetDevice(dev[gidx]));
        int sms = 148;
        AGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev[gidx]));
        const int64_t cols = c_lo[gidx + 1] - c_lo[gidx];
        const int k = pick_k(cols, sms);
        ks[gidx] = k;
        const int64_t n_stripes = (cols + 32 * k - 1) / (32 * k);
        SwLongWorkspace &w = *ws[gidx];
        const int64_t need = 2 * lb + n_stripes + 2 + 1;
        if (need > w.cap) {
            if (w.buf) cudaFree(w.buf);
            w.buf = nullptr; w.cap = 0;
            AGX_CUDA(cudaMalloc(&w.buf, (size_t)need * sizeof(int32_t)));
            w.cap = need;
        }
        const int64_t seq_need = cols + lb + 64;
        if (seq_need > w.cap_seq) {
            if (w.seq) cudaFree(w.seq);
            w.seq = nullptr; w.cap_seq = 0;
            AGX_CUDA(cudaMalloc(&w.seq, (size_t)seq_need));
            w.cap_seq = seq_need;
        }
        AGX_CUDA(cudaMemcpyAsync(w.seq, a + c_lo[gidx], (size_t)cols, cudaMemcpyHostToDevice, st[gidx]));
        AGX_CUDA(cudaMemcpyAsync(w.seq + cols, b, (size_t)lb, cudaMemcpyHostToDevice, st[gidx]));
        LongArgs &x = args[gidx];
        x.a = w.seq; x.la = (int32_t)cols; x.b = w.seq + cols; x.lb = (int32_t)lb;
        x.bnd = w.buf;
        x.progress = w.buf + 2 * lb;
        x.best = w.buf + 2 * lb + n_stripes + 2;
        x.next_bnd = nullptr; x.next_progress = nullptr;
        x.first_gpu = (gidx == 0);
        x.rb = pick_rb();
        x.sc = sc;
        AGX_CUDA(cudaMemsetAsync(x.progress, 0, (size_t)(n_stripes + 3) * sizeof(int32_t), st[gidx]));
    }
    for (int gidx = 0; gidx + 1 < n_dev; ++gidx) {
        args[gidx].next_bnd = args[gidx + 1].bnd;
        args[gidx].next_progress = args[gidx + 1].progress;
    }
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
