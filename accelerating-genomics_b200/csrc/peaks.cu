// peaks.cu -- instruction-throughput microbenchmark: the roofline denominators that
// MEASURED_PEAKS.json does not carry (INT32 / DPX / FP32 lane-ops per clock per SM at the clocks
// the GPU actually holds under this load).  SURVEY.md section 7 step 4.
//
// Each test runs NCHAIN independent dependency chains per thread, fully unrolled, on every SM with
// enough resident warps to saturate the pipe, and reports
//     lane-ops / clk / SM   (from clock64 deltas inside the kernel)
//     T lane-ops / s        (from CUDA events: what the chip sustains at its real clock)
// Output: one JSON object per line on stdout.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e = (x);                                                       \
        if (e != cudaSuccess) {                                                    \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

// ---- SM clock / throttle reasons while the kernels run: NVML through dlopen (no link-time dependency) ----
#include <dlfcn.h>
#include <atomic>
#include <thread>
#include <vector>
#include <algorithm>
struct Nvml {
    void *h = nullptr, *dev = nullptr;
    int (*clock)(void *, int, unsigned *) = nullptr;
    int (*reasons)(void *, unsigned long long *) = nullptr;
    std::atomic<bool> stop{false};
    std::vector<unsigned> mhz;
    unsigned long long seen = 0;
    std::thread th;
    bool open(int index)
    {
        h = dlopen("libnvidia-ml.so.1", RTLD_NOW);
        if (!h) return false;
        auto init = (int (*)())dlsym(h, "nvmlInit_v2");
        auto get = (int (*)(unsigned, void **))dlsym(h, "nvmlDeviceGetHandleByIndex_v2");
        clock = (int (*)(void *, int, unsigned *))dlsym(h, "nvmlDeviceGetClockInfo");
        reasons = (int (*)(void *, unsigned long long *))dlsym(h, "nvmlDeviceGetCurrentClocksEventReasons");
        if (!reasons) reasons = (int (*)(void *, unsigned long long *))dlsym(h, "nvmlDeviceGetCurrentClocksThrottleReasons");
        if (!init || !get || !clock || init() != 0 || get((unsigned)index, &dev) != 0) return false;
        return true;
    }
    void start()
    {
        if (!dev) return;
        stop = false;
        th = std::thread([this] {
            while (!stop.load()) {
                unsigned v = 0;
                if (clock(dev, 1 /* NVML_CLOCK_SM */, &v) == 0) mhz.push_back(v);
                unsigned long long r = 0;
                if (reasons && reasons(dev, &r) == 0) seen |= r;
                std::this_thread::sleep_for(std::chrono::milliseconds(5));
            }
        });
    }
    void finish()
    {
        if (!dev) { printf("{\"clocks\": null}\n"); return; }
        stop = true;
        th.join();
        std::sort(mhz.begin(), mhz.end());
        const unsigned med = mhz.empty() ? 0 : mhz[mhz.size() / 2];
        printf("{\"clocks\": {\"sm_mhz_median\": %u, \"sm_mhz_min\": %u, \"sm_mhz_max\": %u, \"samples\": %zu, "
               "\"reasons_mask\": \"0x%llx\", \"hw_slowdown\": %d, \"hw_thermal_slowdown\": %d, \"sw_thermal_slowdown\": %d, "
               "\"sw_power_cap\": %d}}\n",
               med, mhz.empty() ? 0 : mhz.front(), mhz.empty() ? 0 : mhz.back(), mhz.size(), seen, (int)((seen >> 3) & 1),
               (int)((seen >> 6) & 1), (int)((seen >> 5) & 1), (int)((seen >> 2) & 1));
    }
};

constexpr int NCHAIN = 8;
constexpr int UNROLL = 8;   // ops per chain per loop iteration

enum Op {
    OP_VIADDMNMX_S16X2, OP_VIMNMX3_S16X2_RELU, OP_VIADD_16X2, OP_VIMNMX_S16X2, OP_PRMT,
    OP_VIADDMNMX_S32, OP_VIMNMX3_S32_RELU, OP_IADD3, OP_LOP3, OP_IMAD,
    OP_FFMA, OP_FMUL, OP_FADD, OP_FSEL,
    OP_MIX_DPX_IMAD, OP_MIX_DPX_FFMA, OP_MIX_FFMA_SEL, OP_SWCELL, OP_SHFL,
    OP_MIX_VIADD_VIADDMNMX, OP_MIX_VIADDMNMX_VIMNMX3, OP_MIX_PRMT_VIADDMNMX, OP_MIX_VIADD_PRMT, OP_MIX_VIADD_IMAD,
    OP_MIX_VIMNMX3_IMAD, OP_MIX_PRMT_IMAD, OP_MIX_LOP3_VIADDMNMX, OP_MIX_VIADD_FFMA, OP_SWCELL_FULL,
    OP_HMNMX2, OP_MIX_HMNMX2_VIADDMNMX, OP_MIX_HMNMX2_VIADD,
    OP_IDP4A, OP_MIX_IDP4A_VIADDMNMX, OP_MIX_IDP4A_IMAD, OP_SWCELL_S32, OP_SWCELL_S32_DP4A, OP_COUNT
};
static const char *op_name[OP_COUNT] = {
    "VIADDMNMX.S16x2", "VIMNMX3.S16x2.RELU", "VIADD.16x2", "VIMNMX.S16x2", "PRMT",
    "VIADDMNMX.S32", "VIMNMX3.S32.RELU", "IADD3", "LOP3", "IMAD",
    "FFMA", "FMUL", "FADD", "FSEL(ISETP+SEL)",
    "mix 1 VIADDMNMX.S16x2 : 1 IMAD", "mix 1 VIADDMNMX.S16x2 : 1 FFMA", "mix 3 FFMA : 1 ISETP+FSEL",
    "SW s16x2 cell w/o PRMT (6 ops)", "SHFL.UP",
    "mix VIADD.16x2 : VIADDMNMX.S16x2", "mix VIADDMNMX.S16x2 : VIMNMX3.S16x2", "mix PRMT : VIADDMNMX.S16x2",
    "mix VIADD.16x2 : PRMT", "mix VIADD.16x2 : IMAD", "mix VIMNMX3.S16x2 : IMAD", "mix PRMT : IMAD",
    "mix LOP3 : VIADDMNMX.S16x2", "mix VIADD.16x2 : FFMA", "SW s16x2 cell with PRMT (7 ops)",
    "HMNMX2 (max.f16x2)", "mix HMNMX2 : VIADDMNMX.S16x2", "mix HMNMX2 : VIADD.16x2",
    "IDP.4A (dp4a.s32.s32)", "mix IDP.4A : VIADDMNMX.S32", "mix IDP.4A : IMAD",
    "SW s32 coded cell (PRMT + 2 IMAD + 3 DPX, 6 ops)", "SW s32 cell, dp4a substitution (1 IDP + 1 IMAD + 3 DPX, 5 ops)"};
// lane-ops counted per chain step
static const double op_count[OP_COUNT] = {1, 2, 1, 2, 1, 1, 2, 1, 1, 1, 1, 1, 1, 2, 2, 2, 5, 6, 1,
                                              2, 2, 2, 2, 2, 2, 2, 2, 2, 7, 2, 2, 2,
                                              1, 2, 2, 6, 5};

__device__ __forceinline__ void ffma(uint32_t &x, uint32_t a, uint32_t b)
{
    float fx = __uint_as_float(x);
    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(fx) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)));
    x = __float_as_uint(fx);
}

template <int OP>
__device__ __forceinline__ void step(uint32_t &x, uint32_t &y, uint32_t a, uint32_t b, uint32_t c)
{
    if constexpr (OP == OP_VIADDMNMX_S16X2) x = __viaddmax_s16x2(x, a, b);
    else if constexpr (OP == OP_VIMNMX3_S16X2_RELU) x = __vimax3_s16x2_relu(x, a, b) ^ c;
    else if constexpr (OP == OP_VIADD_16X2) x = __vadd2(x, a);
    else if constexpr (OP == OP_VIMNMX_S16X2) x = __vmaxs2(x, a) ^ c;
    else if constexpr (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    else if constexpr (OP == OP_VIADDMNMX_S32) x = __viaddmax_s32(x, a, b);
    else if constexpr (OP == OP_VIMNMX3_S32_RELU) x = __vimax3_s32_relu(x, a, b) ^ c;
    else if constexpr (OP == OP_IADD3) asm volatile("{.reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2;}" : "+r"(x) : "r"(a), "r"(b));
    else if constexpr (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
    else if constexpr (OP == OP_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    else if constexpr (OP == OP_FFMA) ffma(x, a, b);
    else if constexpr (OP == OP_FMUL) {
        float fx = __uint_as_float(x);
        asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(fx) : "f"(__uint_as_float(a)));
        x = __float_as_uint(fx);
    } else if constexpr (OP == OP_FADD) {
        float fx = __uint_as_float(x);
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(fx) : "f"(__uint_as_float(a)));
        x = __float_as_uint(fx);
    }
    else if constexpr (OP == OP_FSEL) {
        asm volatile("{.reg .pred p; setp.eq.s32 p, %0, %1; selp.b32 %0, %2, %0, p;}" : "+r"(x) : "r"(a), "r"(b));
    } else if constexpr (OP == OP_MIX_DPX_IMAD) {
        x = __viaddmax_s16x2(x, a, b);
        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y) : "r"(a), "r"(b));
    } else if constexpr (OP == OP_MIX_DPX_FFMA) {
        x = __viaddmax_s16x2(x, a, b);
        ffma(y, a, b);
    } else if constexpr (OP == OP_MIX_FFMA_SEL) {
        ffma(x, a, b);
        ffma(y, a, b);
        ffma(x, b, a);
        asm volatile("{.reg .pred p; setp.eq.s32 p, %0, %1; selp.b32 %0, %2, %0, p;}" : "+r"(y) : "r"(c), "r"(b));
    } else if constexpr (OP == OP_SWCELL) {
        // the instruction mix of one sw_duo_kernel cell (x = E chain, y = H/G of the previous row)
        uint32_t t;
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(a), "r"(b), "r"(c));
        const uint32_t d = __vadd2(y, t);
        x = __viaddmax_s16x2(x, a, y);
        const uint32_t f = __viaddmax_s16x2(y, a, d);
        const uint32_t h = __vimax3_s16x2_relu(x, f, d);
        y = __vadd2(h, b);
        x = __vmaxs2(x, y);   // stands in for the running max (0.5 / cell in the real kernel)
    } else if constexpr (OP == OP_SHFL) x = __shfl_up_sync(0xffffffffu, x, 1);
    else if constexpr (OP == OP_MIX_VIADD_VIADDMNMX) { x = __vadd2(x, a); y = __viaddmax_s16x2(y, a, b); }
    else if constexpr (OP == OP_MIX_VIADDMNMX_VIMNMX3) { x = __viaddmax_s16x2(x, a, b); y = __vimax3_s16x2(y, x, c); }
    else if constexpr (OP == OP_MIX_PRMT_VIADDMNMX) {
        asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
        y = __viaddmax_s16x2(y, a, b);
    } else if constexpr (OP == OP_MIX_VIADD_PRMT) {
        x = __vadd2(x, a);
        asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(y) : "r"(a), "r"(b));
    } else if constexpr (OP == OP_MIX_VIADD_IMAD) {
        x = __vadd2(x, a);
        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y) : "r"(a), "r"(b));
    } else if constexpr (OP == OP_MIX_VIMNMX3_IMAD) {
        x = __vimax3_s16x2(x, y, c);
        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y) : "r"(a), "r"(b));
    } else if constexpr (OP == OP_MIX_PRMT_IMAD) {
        asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y) : "r"(a), "r"(b));
    } else if constexpr (OP == OP_MIX_LOP3_VIADDMNMX) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
        y = __viaddmax_s16x2(y, a, b);
    } else if constexpr (OP == OP_MIX_VIADD_FFMA) { x = __vadd2(x, a); ffma(y, a, b); }
    else if constexpr (OP == OP_HMNMX2) {
        // operands change every step, so ptxas cannot fold max(max(x, a), a)
        asm volatile("max.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
        asm volatile("min.f16x2 %0, %0, %1;" : "+r"(y) : "r"(x));
    } else if constexpr (OP == OP_MIX_HMNMX2_VIADDMNMX) {
        y = __viaddmax_s16x2(y, a, b);
        asm volatile("max.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
    } else if constexpr (OP == OP_MIX_HMNMX2_VIADD) {
        y = __vadd2(y, a);
        asm volatile("max.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
    }
    else if constexpr (OP == OP_IDP4A) x = (uint32_t)__dp4a((int)a, (int)b, (int)x);
    else if constexpr (OP == OP_MIX_IDP4A_VIADDMNMX) { x = (uint32_t)__dp4a((int)a, (int)b, (int)x); y = __viaddmax_s32(y, a, b); }
    else if constexpr (OP == OP_MIX_IDP4A_IMAD) {
        x = (uint32_t)__dp4a((int)a, (int)b, (int)x);
        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y) : "r"(a), "r"(b));
    } else if constexpr (OP == OP_SWCELL_S32) {
        // one symbol-coded s32 cell of sw_long.cu: x = E chain, y = H + goe of the previous row
        uint32_t t, d, g;
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(a), "r"(b), "r"(c));
        asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(y), "r"(1 | (c & 0)), "r"(t));
        x = __viaddmax_s32(x, a, y);
        const uint32_t f = __viaddmax_s32(y, a, d);
        const uint32_t h = __vimax3_s32_relu(x, f, d);
        asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(g) : "r"(h), "r"(1 | (c & 0)), "r"(b));
        y = g;
    } else if constexpr (OP == OP_SWCELL_S32_DP4A) {
        // the same cell with the substitution score added by one dp4a (byte-select by a one-hot multiplier)
        uint32_t g;
        const uint32_t d = (uint32_t)__dp4a((int)a, (int)(0x00000100u | (c & 0)), (int)y);
        x = __viaddmax_s32(x, a, y);
        const uint32_t f = __viaddmax_s32(y, a, d);
        const uint32_t h = __vimax3_s32_relu(x, f, d);
        asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(g) : "r"(h), "r"(1 | (c & 0)), "r"(b));
        y = g;
    }
    else if constexpr (OP == OP_SWCELL_FULL) {
        uint32_t t;
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(x), "r"(b), "r"(c));
        const uint32_t d = __vadd2(y, t);
        x = __viaddmax_s16x2(x, a, y);
        const uint32_t f = __viaddmax_s16x2(y, a, d);
        const uint32_t h = __vimax3_s16x2_relu(x, f, d);
        y = __vadd2(h, b);
        x = __vmaxs2(x, y);
    }
}

template <int OP>
__global__ void __launch_bounds__(256) bench_kernel(uint32_t *out, long long *cycles, uint32_t seed, int iters)
{
    uint32_t x[NCHAIN], y[NCHAIN];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) { x[i] = seed * (tid + i + 1); y[i] = seed + tid * 7 + i; }
    const uint32_t a = seed | 0x00010001u, b = seed ^ 0x3f800000u, c = (seed >> 3) & 0x3210u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < NCHAIN; ++i) step<OP>(x[i], y[i], a, b, c);
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) acc ^= x[i] ^ y[i];
    out[tid] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, uint32_t *d_out, long long *d_cyc, int blocks_per_sm)
{
    const int threads = 256, iters = 2048;
    const int blocks = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) bench_kernel<OP><<<blocks, threads>>>(d_out, d_cyc, 12345u + w, iters);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    double best_cyc = 0;
    long long *h_cyc = (long long *)malloc(sizeof(long long) * blocks);
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        bench_kernel<OP><<<blocks, threads>>>(d_out, d_cyc, 777u + rep, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ms) {
            best_ms = ms;
            CK(cudaMemcpy(h_cyc, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
            double s = 0;
            for (int i = 0; i < blocks; ++i) s += (double)h_cyc[i];
            best_cyc = s / blocks;
        }
    }
    free(h_cyc);
    const double ops_per_thread = (double)iters * UNROLL * NCHAIN * op_count[OP];
    const double lane_ops = ops_per_thread * threads * blocks;
    // all blocks of an SM are co-resident (blocks_per_sm * 256 threads <= 2048), so an SM's work
    // is blocks_per_sm blocks during ~best_cyc cycles
    // (clock64 is a per-SM cycle counter: mean block span; the SM clock itself is in the NVML "clocks" record)
    const double per_clk_sm = ops_per_thread * threads * blocks_per_sm / best_cyc;
    printf("{\"op\": \"%s\", \"lane_ops_per_clk_per_sm\": %.2f, \"tera_lane_ops_per_s\": %.3f, "
           "\"ms\": %.4f, \"warps_per_sm\": %d}\n",
           op_name[OP], per_clk_sm, lane_ops / (best_ms * 1e-3) / 1e12, best_ms, blocks_per_sm * threads / 32);
    fflush(stdout);
}

// ---- FP32 with realistic operand patterns ---------------------------------------------------------
// (a) FFMA whose three sources are three DIFFERENT registers (the ordinary peak test reuses a and b)
// (b) the PairHMM cell of hmm_stream_kernel: K rows, per-row coefficients in registers, the same
//     dependency structure (independent M terms, X chain down the rows), no loads, no shuffles
template <int MODE, int K>
__global__ void __launch_bounds__(256) fp_pattern_kernel(float *out, long long *cycles, float seed, int iters)
{
    float ca[K], cbx[K], cby[K], ccx[K], cg[K], pr[K], M[K], X[K], Y[K];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        ca[j] = 0.9f + seed * (j + 1); cbx[j] = 1e-4f * (j + 1) + seed; cby[j] = 2e-4f * (j + 2) + seed;
        ccx[j] = 0.1f + seed * j; cg[j] = 0.1f + seed * (j + 3); pr[j] = 0.99f - seed * j;
        M[j] = seed * tid + j; X[j] = seed + j; Y[j] = 1.f + seed * j;
    }
    float upM0 = seed, upX0 = seed * 2, dM0 = seed * 3, dX0 = seed * 4, dY0 = seed * 5;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 2
    for (int it = 0; it < iters; ++it) {
        if constexpr (MODE == 0) {
#pragma unroll
            for (int j = 0; j < K; ++j) {
                M[j] = fmaf(ca[j], cbx[j], M[j]);
                X[j] = fmaf(cby[j], ccx[j], X[j]);
                Y[j] = fmaf(cg[j], pr[j], Y[j]);
            }
        } else {
            float upM = upM0, upX = upX0, dM = dM0, dX = dX0, dY = dY0;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const float oM = M[j], oX = X[j], oY = Y[j];
                float vv = cby[j] * dY;
                vv = fmaf(cbx[j], dX, vv);
                vv = fmaf(ca[j], dM, vv);
                const float mn = pr[j] * vv;
                const float xn = fmaf(ccx[j], upX, upM);
                const float yn = fmaf(cg[j], oY, oM);
                dM = oM; dX = oX; dY = oY;
                upM = mn; upX = xn;
                M[j] = mn; X[j] = xn; Y[j] = yn;
            }
            upM0 = M[K - 1]; upX0 = X[K - 1]; dM0 = upM0; dX0 = upX0; dY0 = Y[K - 1];
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) acc += M[j] + X[j] + Y[j];
    out[tid] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int K> void run_fp_pattern(int sms, uint32_t *d_out, long long *d_cyc, int blocks_per_sm, const char *name)
{
    const int threads = 256, iters = 4096;
    const int blocks = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) fp_pattern_kernel<MODE, K><<<blocks, threads>>>((float *)d_out, d_cyc, 1e-3f, iters);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        fp_pattern_kernel<MODE, K><<<blocks, threads>>>((float *)d_out, d_cyc, 1e-3f + rep * 1e-6f, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ms) best_ms = ms;
    }
    const double per_thread = (double)iters * K * (MODE == 0 ? 3 : 6);
    const double lane_ops = per_thread * threads * blocks;
    printf("{\"op\": \"%s\", \"tera_lane_ops_per_s\": %.3f, \"ms\": %.4f, \"warps_per_sm\": %d}\n", name,
           lane_ops / (best_ms * 1e-3) / 1e12, best_ms, blocks_per_sm * threads / 32);
    fflush(stdout);
}


// ---- packed FP32 (sm_100a FFMA2 / FMUL2 / FADD2 = PTX fma.rn.f32x2 ...) ---------------------------
// One instruction, two FP32 lanes per thread.  Questions answered here: does the packed form raise
// the FP32 lane throughput, and does it free issue slots for the other pipes?
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}

// MODE 0: FFMA2 only; 1: FFMA2 with a scalar (broadcast) multiplier; 2: 1 FFMA2 : 1 LOP3;
// 3: 1 FFMA2 : 1 FFMA; 4: 1 FFMA2 : 1 SHFL.UP (per 4); 5: 1 FFMA2 : 2 LOP3
template <int MODE>
__global__ void __launch_bounds__(256) fp2_kernel(float *out, long long *cycles, float seed, int iters)
{
    float2 x[NCHAIN];
    uint32_t y[NCHAIN];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) { x[i] = make_float2(seed * (tid + i), seed + i); y[i] = tid * 7 + i; }
    const float2 a = make_float2(1.f + seed, 1.f - seed), b = make_float2(seed, -seed);
    const float sc = 1.f + 2 * seed;
    const uint32_t ia = __float_as_uint(seed) | 0x10001u, ib = __float_as_uint(seed) ^ 0x3f800000u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < NCHAIN; ++i) {
                if constexpr (MODE == 1) x[i] = ffma2(x[i], make_float2(sc, sc), b);
                else x[i] = ffma2(x[i], a, b);
                if constexpr (MODE == 2 || MODE == 5)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(ia), "r"(ib));
                if constexpr (MODE == 5)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(ib), "r"(ia));
                if constexpr (MODE == 3) ffma(y[i], ia, ib);
                if constexpr (MODE == 4) { if ((i & 3) == 0) y[i] = __shfl_up_sync(0xffffffffu, y[i], 1); }
            }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) acc += x[i].x + x[i].y + __uint_as_float(y[i]);
    out[tid] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> void run_fp2(int sms, uint32_t *d_out, long long *d_cyc, int blocks_per_sm, const char *name,
                                 double fp_lane_ops, double other_ops)
{
    const int threads = 256, iters = 2048;
    const int blocks = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) fp2_kernel<MODE><<<blocks, threads>>>((float *)d_out, d_cyc, 1e-3f, iters);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        fp2_kernel<MODE><<<blocks, threads>>>((float *)d_out, d_cyc, 1e-3f + rep * 1e-6f, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ms) best_ms = ms;
    }
    const double steps = (double)iters * UNROLL * NCHAIN * threads * blocks;
    printf("{\"op\": \"%s\", \"fp32_tera_lane_ops_per_s\": %.3f, \"other_tera_lane_ops_per_s\": %.3f, "
           "\"tera_warp_instr_x32_per_s\": %.3f, \"ms\": %.4f, \"warps_per_sm\": %d}\n",
           name, steps * fp_lane_ops / (best_ms * 1e-3) / 1e12, steps * other_ops / (best_ms * 1e-3) / 1e12,
           steps * (1 + other_ops) / (best_ms * 1e-3) / 1e12, best_ms, blocks_per_sm * threads / 32);
    fflush(stdout);
}

// PairHMM cell with two haplotype columns per lane packed in f32x2 (per-row coefficients are scalars,
// broadcast by FFMA2's .F32 operand form): 5 packed instructions + 0 unpacked per 2 cells when the prior
// is already a pair, MODE 1: prior multiply unpacked (2 FMUL) as in a per-symbol table lookup
template <int MODE, int K>
__global__ void __launch_bounds__(256) fp2_pattern_kernel(float *out, long long *cycles, float seed, int iters)
{
    float ca[K], cbx[K], cby[K], ccx[K], cg[K];
    float2 pr[K], M[K], X[K], Y[K];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        ca[j] = 0.9f + seed * (j + 1); cbx[j] = 1e-4f * (j + 1) + seed; cby[j] = 2e-4f * (j + 2) + seed;
        ccx[j] = 0.1f + seed * j; cg[j] = 0.1f + seed * (j + 3); pr[j] = make_float2(0.99f - seed * j, 0.01f + seed * j);
        M[j] = make_float2(seed * tid + j, seed); X[j] = make_float2(seed + j, seed * 2); Y[j] = make_float2(1.f + seed * j, 1.f);
    }
    float2 upM0 = make_float2(seed, seed), upX0 = upM0, dM0 = upM0, dX0 = upM0, dY0 = upM0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 2
    for (int it = 0; it < iters; ++it) {
        float2 upM = upM0, upX = upX0, dM = dM0, dX = dX0, dY = dY0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const float2 oM = M[j], oX = X[j], oY = Y[j];
            float2 vv = fmul2(make_float2(cby[j], cby[j]), dY);
            vv = ffma2(make_float2(cbx[j], cbx[j]), dX, vv);
            vv = ffma2(make_float2(ca[j], ca[j]), dM, vv);
            float2 mn;
            if constexpr (MODE == 0) mn = fmul2(pr[j], vv);
            else { mn.x = pr[j].x * vv.x; mn.y = pr[j].y * vv.y; }
            const float2 xn = ffma2(make_float2(ccx[j], ccx[j]), upX, upM);
            const float2 yn = ffma2(make_float2(cg[j], cg[j]), oY, oM);
            dM = oM; dX = oX; dY = oY;
            upM = mn; upX = xn;
            M[j] = mn; X[j] = xn; Y[j] = yn;
        }
        upM0 = M[K - 1]; upX0 = X[K - 1]; dM0 = upM0; dX0 = upX0; dY0 = Y[K - 1];
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) acc += M[j].x + X[j].x + Y[j].x + M[j].y + X[j].y + Y[j].y;
    out[tid] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int K> void run_fp2_pattern(int sms, uint32_t *d_out, long long *d_cyc, int blocks_per_sm, const char *name)
{
    const int threads = 256, iters = 4096;
    const int blocks = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) fp2_pattern_kernel<MODE, K><<<blocks, threads>>>((float *)d_out, d_cyc, 1e-3f, iters);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        fp2_pattern_kernel<MODE, K><<<blocks, threads>>>((float *)d_out, d_cyc, 1e-3f + rep * 1e-6f, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ms) best_ms = ms;
    }
    const double cells = (double)iters * K * 2 * threads * blocks;
    printf("{\"op\": \"%s\", \"tera_cells_per_s\": %.3f, \"fp32_tera_lane_ops_per_s\": %.3f, \"ms\": %.4f, \"warps_per_sm\": %d}\n",
           name, cells / (best_ms * 1e-3) / 1e12, cells * 6 / (best_ms * 1e-3) / 1e12, best_ms, blocks_per_sm * threads / 32);
    fflush(stdout);
}


// ---- the whole inner loop of the PairHMM stream kernel, not just its FP32 part ----------------------
// One warp per block as in hmm_stream_kernel: per step a symbol prefetch (LDG.U8), the prior lookup
// (LDS.128 from a per-warp table), the three boundary shuffles, the lane-0 selects, K cells, the running
// sum and the haplotype counter.  PACKED = 0: one read per warp, scalar FP32 (what ships today);
// PACKED = 1: two reads per warp, every FP32 instruction in its f32x2 form (FFMA2 / FMUL2).
template <int PACKED, int K>
__global__ void __launch_bounds__(32) hmm_loop_kernel(float *out, const uint8_t *codes, float seed, int steps)
{
    constexpr int V = PACKED ? 2 : 1;
    constexpr int CH = (K * V + 3) / 4;
    __shared__ float4 tab[6][CH][32];
    const int t = threadIdx.x;
    for (int sgn = 0; sgn < 6; ++sgn)
        for (int c = 0; c < CH; ++c) tab[sgn][c][t] = make_float4(0.9f + seed * sgn, 0.01f + seed * c, 0.9f, 0.02f + seed * t);
    __syncwarp();
    float2 ca[K], cbx[K], cby[K], ccx[K], cg[K], M[K], X[K], Y[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        // coefficients come from memory so that they have to live in registers, as in the real kernel
        const float2 *cf = reinterpret_cast<const float2 *>(out) + (size_t)(blockIdx.x * 32 + t) * 0 + j * 5;
        ca[j] = cf[0]; cbx[j] = cf[1]; cby[j] = cf[2]; ccx[j] = cf[3]; cg[j] = cf[4];
        M[j] = make_float2(seed * t + j, seed); X[j] = make_float2(seed + j, seed * 2); Y[j] = make_float2(1.f + seed * j, 1.f);
    }
    float2 pdM = make_float2(0.f, 0.f), pdX = pdM, pdY = pdM, bM = pdM, bX = pdM, bY = pdM, acc = pdM;
    const float2 qi_last = make_float2(0.5f + seed, 0.5f - seed);
    const float init = 1.f + seed;
    const uint8_t *cp = codes + (blockIdx.x & 1023) * 64;
    uint32_t code_next = *cp;
    int rem = steps + 5;
    const float4 *tab_lane = &tab[0][0][t];
    const bool lane0 = (t == 0);
#pragma unroll 2
    for (int s = 0; s < steps; ++s) {
        if (rem == 0) { rem = steps; acc = make_float2(0.f, 0.f); cp = codes; }
        const uint32_t code = code_next;
        cp += 1;
        code_next = *cp;
        float4 pr4[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) pr4[c] = tab_lane[code * (CH * 32) + c * 32];
        float2 upM, upX, upY;
        upM.x = __shfl_up_sync(0xffffffffu, bM.x, 1); upX.x = __shfl_up_sync(0xffffffffu, bX.x, 1); upY.x = __shfl_up_sync(0xffffffffu, bY.x, 1);
        if (PACKED) { upM.y = __shfl_up_sync(0xffffffffu, bM.y, 1); upX.y = __shfl_up_sync(0xffffffffu, bX.y, 1); upY.y = __shfl_up_sync(0xffffffffu, bY.y, 1); }
        else { upM.y = upX.y = upY.y = 0.f; }
        if (lane0) { upM = make_float2(0.f, 0.f); upX = upM; upY = make_float2(init, init); }
        float2 dM = pdM, dX = pdX, dY = pdY;
        pdM = upM; pdX = upX; pdY = upY;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const float2 oM = M[j], oX = X[j], oY = Y[j];
            float2 mn, xn, yn;
            if (PACKED) {
                const float4 p4 = pr4[j / 2];
                const float2 pr = (j & 1) ? make_float2(p4.z, p4.w) : make_float2(p4.x, p4.y);
                float2 vv = fmul2(cby[j], dY);
                vv = ffma2(cbx[j], dX, vv);
                vv = ffma2(ca[j], dM, vv);
                mn = fmul2(pr, vv);
                xn = ffma2(ccx[j], upX, upM);
                yn = ffma2(cg[j], oY, oM);
            } else {
                const float4 p4 = pr4[j / 4];
                const float pr = (j % 4 == 0) ? p4.x : (j % 4 == 1) ? p4.y : (j % 4 == 2) ? p4.z : p4.w;
                float vv = cby[j].x * dY.x;
                vv = fmaf(cbx[j].x, dX.x, vv);
                vv = fmaf(ca[j].x, dM.x, vv);
                mn.x = pr * vv; mn.y = 0.f;
                xn.x = fmaf(ccx[j].x, upX.x, upM.x); xn.y = 0.f;
                yn.x = fmaf(cg[j].x, oY.x, oM.x); yn.y = 0.f;
            }
            dM = oM; dX = oX; dY = oY;
            upM = mn; upX = xn;
            M[j] = mn; X[j] = xn; Y[j] = yn;
        }
        bM = M[K - 1]; bX = X[K - 1]; bY = Y[K - 1];
        if (PACKED) acc = ffma2(qi_last, bX, ffma2(make_float2(1.f, 1.f), bM, acc));
        else acc.x += fmaf(qi_last.x, bX.x, bM.x);
        --rem;
    }
    float r = acc.x + acc.y;
#pragma unroll
    for (int j = 0; j < K; ++j) r += M[j].x + M[j].y + Y[j].x + Y[j].y;
    out[blockIdx.x * 32 + t] = r;
}

template <int PACKED, int K> void run_hmm_loop(int sms, int warps_per_sm, const char *name)
{
    const int blocks = sms * warps_per_sm, steps = 20000;
    float *d_out;
    uint8_t *d_codes;
    CK(cudaMalloc(&d_out, sizeof(float) * blocks * 32));
    CK(cudaMalloc(&d_codes, 1 << 20));
    {
        uint8_t *h = (uint8_t *)malloc(1 << 20);
        for (int i = 0; i < (1 << 20); ++i) h[i] = (uint8_t)((i * 2654435761u >> 13) % 5);
        CK(cudaMemcpy(d_codes, h, 1 << 20, cudaMemcpyHostToDevice));
        free(h);
    }
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hmm_loop_kernel<PACKED, K>, 32, 0));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, hmm_loop_kernel<PACKED, K>));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) hmm_loop_kernel<PACKED, K><<<blocks, 32>>>(d_out, d_codes, 1e-3f, steps);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        hmm_loop_kernel<PACKED, K><<<blocks, 32>>>(d_out, d_codes, 1e-3f + rep * 1e-6f, steps);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ms) best_ms = ms;
    }
    const double cells = (double)steps * K * (PACKED ? 2 : 1) * 32 * blocks;
    printf("{\"op\": \"%s\", \"tera_cells_per_s\": %.3f, \"ms\": %.4f, \"warps_per_sm_launched\": %d, "
           "\"max_warps_per_sm\": %d, \"registers\": %d}\n",
           name, cells / (best_ms * 1e-3) / 1e12, best_ms, warps_per_sm, occ, fa.numRegs);
    fflush(stdout);
    CK(cudaFree(d_out));
    CK(cudaFree(d_codes));
}

// ---- issue rate of ONE warp (and of 2, 4, 8) per SM sub-partition ---------------------------------------
// CH independent chains per thread, every operand in its own register (no operand-reuse cache, no uniform
// register): what a lone warp of the long-alignment kernel can get out of a pipe.
enum IssueOp { IS_VIADDMNMX, IS_VIMNMX3, IS_IMAD, IS_IDP4A, IS_CELL, IS_COUNT };
template <int IOP, int CH>
__global__ void __launch_bounds__(256) issue_kernel(uint32_t *out, uint32_t seed, int iters)
{
    int32_t x[CH], y[CH], z[CH];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < CH; ++i) { x[i] = (int32_t)(seed * (tid + i + 1)); y[i] = (int32_t)(seed + tid * 7 + i); z[i] = (int32_t)(seed ^ (tid + 3 * i)); }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if constexpr (IOP == IS_VIADDMNMX) x[i] = __viaddmax_s32(x[i], y[i], z[i]);
                else if constexpr (IOP == IS_VIMNMX3) x[i] = __vimax3_s32_relu(x[i], y[i], z[i]) - 1;
                else if constexpr (IOP == IS_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(z[i]));
                else if constexpr (IOP == IS_IDP4A) x[i] = __dp4a(y[i], z[i], x[i]);
                else {
                    // the lean symbol-coded cell: x = E chain, y = H + goe above, z = F
                    const int32_t d = __dp4a(z[i], 0x100, y[i]);
                    x[i] = __viaddmax_s32(x[i], -1, y[i]);
                    z[i] = __viaddmax_s32(z[i], -1, y[i]);
                    int32_t g;
                    const int32_t h = __vimax3_s32_relu(x[i], z[i], d);
                    asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(g) : "r"(h), "r"((int32_t)(seed | 1)), "r"(-4));
                    y[i] = g;
                }
            }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) acc ^= (uint32_t)(x[i] ^ y[i] ^ z[i]);
    out[tid] = acc;
}
template <int IOP, int CH> void run_issue(int sms, uint32_t *d_out, int warps_per_smsp, const char *name, double ops_per_step)
{
    const int iters = 4096;
    const int threads = warps_per_smsp >= 2 ? 256 : 128;
    const int bps = warps_per_smsp >= 2 ? warps_per_smsp / 2 : 1;
    const int blocks = sms * bps;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    issue_kernel<IOP, CH><<<blocks, threads>>>(d_out, 3u, iters);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        issue_kernel<IOP, CH><<<blocks, threads>>>(d_out, 5u + rep, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double warp_instr_per_smsp = (double)iters * 8 * CH * ops_per_step * warps_per_smsp;
    printf("{\"issue_test\": \"%s\", \"chains\": %d, \"warps_per_smsp\": %d, \"ms\": %.4f, "
           "\"clk_per_warp_instr_at_1965MHz\": %.3f}\n",
           name, CH, warps_per_smsp, best, best * 1e-3 * 1.965e9 / warp_instr_per_smsp);
    fflush(stdout);
}
template <int IOP> void run_issue_all(int sms, uint32_t *d_out, const char *name, double ops)
{
    for (int w : {1, 2, 4, 8}) {
        run_issue<IOP, 1>(sms, d_out, w, name, ops);
        run_issue<IOP, 2>(sms, d_out, w, name, ops);
        run_issue<IOP, 4>(sms, d_out, w, name, ops);
        run_issue<IOP, 8>(sms, d_out, w, name, ops);
    }
}

// ---- the R x K tile of the long-alignment kernel in isolation: one warp per SM sub-partition -------------
// Same cell (dp4a substitution, lean chain), same loop-carried structure (the next step's first cells need this
// step's last column through a shuffle).  ORDER picks how the source lists the cells of a step: 0 column by column
// (rows inner), 1 row by row, 2 anti-diagonals, 3 the parallelogram (row i runs i columns behind and wraps into
// the next step).  FLAGS add the per-step costs of the real kernel one at a time: 1 row tables from a shared-memory
// ring, 2 requested one step ahead, 4 lane 0 takes its boundary from the ring (loads + selects), 8 lane 31 stages
// its boundary in shared memory, 16 two running-max accumulators instead of one.
template <int K, int R, int ORDER, int FLAGS>
__global__ void __launch_bounds__(128) tile_kernel(uint32_t *out, int steps, int goe_, int ext_)
{
    __shared__ __align__(16) int32_t tab[4][64 * 8];
    __shared__ __align__(16) int32_t rge[4][64 * 8];
    __shared__ __align__(16) int2 stage[4][64 * 4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int i = lane; i < 64 * 8; i += 32) { tab[wib][i] = (int32_t)(0xfbfbfbfbu ^ (0x04u << (8 * ((i * 7 + lane) & 3)))); rge[wib][i] = goe_; }
    __syncwarp();
    const int32_t goe = goe_, ext = ext_;
    int32_t Gp[K], F[K], acol[(K + 3) / 4];
#pragma unroll
    for (int j = 0; j < K; ++j) { Gp[j] = goe; F[j] = goe; }
#pragma unroll
    for (int q = 0; q < (K + 3) / 4; ++q) acol[q] = 0x3210 + lane + q;
    int32_t g_out[R], e_out[R], e[R], gleft[R], gdiag[R], tlo[R], thi[R];
    uint32_t sc4[R];
#pragma unroll
    for (int i = 0; i < R; ++i) { g_out[i] = goe; e_out[i] = goe; e[i] = goe; gleft[i] = goe; gdiag[i] = goe; tlo[i] = thi[i] = (int32_t)0xfbfbfbfb; sc4[i] = 0; }
    int32_t best[2] = {goe, goe}, gprev = goe, gcarry = goe;
    auto cell = [&](int i, int j) {
        if ((j & 3) == 0) sc4[i] = __byte_perm((uint32_t)tlo[i], (uint32_t)thi[i], (uint32_t)acol[j >> 2]);
        const int32_t d = __dp4a((int32_t)sc4[i], (int32_t)(1u << (8 * (j & 3))), gdiag[i]);
        const int32_t f = __viaddmax_s32(F[j], ext, Gp[j]);
        e[i] = __viaddmax_s32(e[i], ext, gleft[i]);
        int32_t gn = __vimax3_s32_relu(e[i], f, d);
        asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(gn) : "r"(gn), "r"(ext_ * ext_), "r"(goe));
        gdiag[i] = Gp[j]; Gp[j] = gn; F[j] = f; gleft[i] = gn;
        if ((FLAGS & 16) && (i & 1)) best[1] = max(best[1], gn); else best[0] = max(best[0], gn);
    };
    auto load_tab = [&](int s, int32_t (&lo)[R], int32_t (&hi)[R]) {
        if constexpr (FLAGS & 1) {
            const int sl = (R * (s - lane)) & (64 * R - 1);
#pragma unroll
            for (int i = 0; i < R; ++i) { lo[i] = tab[wib][sl + i]; hi[i] = tab[wib][(sl + i + 256) & 511]; }
        } else {
#pragma unroll
            for (int i = 0; i < R; ++i) { lo[i] = (int32_t)0xfbfbfbfb ^ (s & 4); hi[i] = (int32_t)0xfbfbfbfb; }
        }
    };
    int32_t nlo[R], nhi[R];
    load_tab(0, nlo, nhi);
#pragma unroll 1
    for (int s = 0; s < steps; ++s) {
        int32_t plo[R], phi[R];
        if constexpr (FLAGS & 2) load_tab(s + 1, plo, phi); else load_tab(s, nlo, nhi);
        int32_t bg[R], be[R];
        if constexpr (FLAGS & 4) {
            const int sl0 = (R * s) & (64 * R - 1);
#pragma unroll
            for (int i = 0; i < R; ++i) { bg[i] = rge[wib][sl0 + i]; be[i] = rge[wib][(sl0 + i + 256) & 511]; }
        } else {
#pragma unroll
            for (int i = 0; i < R; ++i) { bg[i] = goe; be[i] = goe; }
        }
        if constexpr (ORDER != 3) {
            int32_t g_in[R];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                g_in[i] = __shfl_up_sync(0xffffffffu, g_out[i], 1);
                e[i] = __shfl_up_sync(0xffffffffu, e_out[i], 1);
                if (lane == 0) { g_in[i] = bg[i]; e[i] = be[i]; }
                gdiag[i] = i == 0 ? gprev : g_in[i - 1];
                gleft[i] = g_in[i];
                tlo[i] = nlo[i]; thi[i] = nhi[i];
            }
            gprev = g_in[R - 1];
            if constexpr (ORDER == 0) {
#pragma unroll
                for (int j = 0; j < K; ++j)
#pragma unroll
                    for (int i = 0; i < R; ++i) cell(i, j);
            } else if constexpr (ORDER == 1) {
#pragma unroll
                for (int i = 0; i < R; ++i)
#pragma unroll
                    for (int j = 0; j < K; ++j) cell(i, j);
            } else {
#pragma unroll
                for (int dg = 0; dg < K + R - 1; ++dg)
#pragma unroll
                    for (int i = 0; i < R; ++i)
                        if (dg - i >= 0 && dg - i < K) cell(i, dg - i);
            }
#pragma unroll
            for (int i = 0; i < R; ++i) { g_out[i] = gleft[i]; e_out[i] = e[i]; }
            if constexpr (FLAGS & 8) {
                if (lane == 31) {
#pragma unroll
                    for (int i = 0; i < R; ++i) stage[wib][((s & 31) * R + i) & 255] = make_int2(g_out[i], e_out[i]);
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < K; ++c) {
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const int j = (c >= i) ? c - i : K + c - i;
                    if (j == 0) {
                        const int32_t gin = (lane == 0) ? bg[i] : g_out[i];
                        const int32_t ein = (lane == 0) ? be[i] : e_out[i];
                        gdiag[i] = gcarry; gcarry = gin; gleft[i] = gin; e[i] = ein; tlo[i] = nlo[i]; thi[i] = nhi[i];
                    }
                    cell(i, j);
                    if (j == K - 1) {
                        g_out[i] = __shfl_up_sync(0xffffffffu, gleft[i], 1);
                        e_out[i] = __shfl_up_sync(0xffffffffu, e[i], 1);
                        if constexpr (FLAGS & 8) { if (lane == 31) stage[wib][((s & 31) * R + i) & 255] = make_int2(gleft[i], e[i]); }
                    }
                }
            }
        }
        if constexpr (FLAGS & 2) {
#pragma unroll
            for (int i = 0; i < R; ++i) { nlo[i] = plo[i]; nhi[i] = phi[i]; }
        }
    }
    uint32_t acc = (uint32_t)(best[0] ^ best[1]);
#pragma unroll
    for (int j = 0; j < K; ++j) acc ^= (uint32_t)(Gp[j] + F[j]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + stage[wib][lane].x;
}
template <int K, int R, int ORDER, int FLAGS> void run_tile(int sms, uint32_t *d_out)
{
    const int steps = 20000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    tile_kernel<K, R, ORDER, FLAGS><<<sms, 128>>>(d_out, steps, -4, -1);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        tile_kernel<K, R, ORDER, FLAGS><<<sms, 128>>>(d_out, steps, -4, -1);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double clk = best * 1e-3 * 1.965e9 / steps;
    printf("{\"tile_test\": 1, \"K\": %d, \"R\": %d, \"order\": %d, \"flags\": %d, \"clk_per_step\": %.1f, \"clk_per_cell\": %.2f}\n", K, R,
           ORDER, FLAGS, clk, clk / (K * R));
    fflush(stdout);
}
template <int K, int R, int ORDER> void run_tile_flags(int sms, uint32_t *d_out)
{
    run_tile<K, R, ORDER, 0>(sms, d_out);
    run_tile<K, R, ORDER, 16>(sms, d_out);
    run_tile<K, R, ORDER, 1>(sms, d_out);
    run_tile<K, R, ORDER, 3>(sms, d_out);
    run_tile<K, R, ORDER, 5>(sms, d_out);
    run_tile<K, R, ORDER, 13>(sms, d_out);
    run_tile<K, R, ORDER, 15>(sms, d_out);
    run_tile<K, R, ORDER, 31>(sms, d_out);
}
template <int K, int R> void run_tile_orders(int sms, uint32_t *d_out)
{
    run_tile_flags<K, R, 0>(sms, d_out);
    run_tile_flags<K, R, 1>(sms, d_out);
    if constexpr (R <= K) run_tile_flags<K, R, 3>(sms, d_out);
}

template <int OP> void run_all(int sms, uint32_t *d_out, long long *d_cyc, int bps)
{
    if constexpr (OP < OP_COUNT) {
        run<OP>(sms, d_out, d_cyc, bps);
        run_all<OP + 1>(sms, d_out, d_cyc, bps);
    }
}

int main(int argc, char **argv)
{
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev));
    const int sms = p.multiProcessorCount;
    const int bps = argc > 1 ? atoi(argv[1]) : 4;   // 4 blocks x 8 warps = 32 warps / SM
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"blocks_per_sm\": %d}\n", p.name, sms,
           p.clockRate, bps);
    uint32_t *d_out;
    long long *d_cyc;
    CK(cudaMalloc(&d_out, sizeof(uint32_t) * sms * bps * 256));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * bps));
    Nvml nv;
    nv.open(dev);
    nv.start();
    struct Fin { Nvml &n; ~Fin() { n.finish(); } } fin{nv};
    if (argc > 2 && !strcmp(argv[2], "tile")) {
        run_tile_orders<7, 2>(sms, d_out);
        run_tile_orders<7, 4>(sms, d_out);
        return 0;
    }
    if (argc > 2 && !strcmp(argv[2], "issue")) {
        run_issue_all<IS_VIADDMNMX>(sms, d_out, "VIADDMNMX.S32 (3 registers)", 1);
        run_issue_all<IS_VIMNMX3>(sms, d_out, "VIMNMX3.S32.RELU + IADD (3 registers)", 2);
        run_issue_all<IS_IMAD>(sms, d_out, "IMAD (3 registers)", 1);
        run_issue_all<IS_IDP4A>(sms, d_out, "IDP.4A (3 registers)", 1);
        run_issue_all<IS_CELL>(sms, d_out, "lean coded cell: IDP + 2 VIADDMNMX + VIMNMX3.RELU + IMAD", 5);
        return 0;
    }
    if (argc > 2 && !strcmp(argv[2], "dp4a")) {
        run<OP_PRMT>(sms, d_out, d_cyc, bps);
        run<OP_IMAD>(sms, d_out, d_cyc, bps);
        run<OP_VIADDMNMX_S32>(sms, d_out, d_cyc, bps);
        run<OP_IDP4A>(sms, d_out, d_cyc, bps);
        run<OP_MIX_IDP4A_VIADDMNMX>(sms, d_out, d_cyc, bps);
        run<OP_MIX_IDP4A_IMAD>(sms, d_out, d_cyc, bps);
        run<OP_SWCELL_S32>(sms, d_out, d_cyc, bps);
        run<OP_SWCELL_S32_DP4A>(sms, d_out, d_cyc, bps);
        return 0;
    }
    const bool only_fp = argc > 2 && !strcmp(argv[2], "fp");
    const bool only_hm = argc > 2 && !strcmp(argv[2], "hm");
    if (only_hm) {
        run<OP_HMNMX2>(sms, d_out, d_cyc, bps);
        run<OP_MIX_HMNMX2_VIADDMNMX>(sms, d_out, d_cyc, bps);
        run<OP_MIX_HMNMX2_VIADD>(sms, d_out, d_cyc, bps);
        return 0;
    }
    if (argc > 2 && !strcmp(argv[2], "loop")) {
        run_hmm_loop<0, 8>(sms, 20, "PairHMM inner loop, scalar FP32, K=8, 20 warps/SM");
        run_hmm_loop<0, 6>(sms, 20, "PairHMM inner loop, scalar FP32, K=6, 20 warps/SM");
        run_hmm_loop<1, 8>(sms, 12, "PairHMM inner loop, two reads f32x2, K=8, 12 warps/SM");
        run_hmm_loop<1, 8>(sms, 8, "PairHMM inner loop, two reads f32x2, K=8, 8 warps/SM");
        run_hmm_loop<1, 6>(sms, 12, "PairHMM inner loop, two reads f32x2, K=6, 12 warps/SM");
        run_hmm_loop<1, 6>(sms, 16, "PairHMM inner loop, two reads f32x2, K=6, 16 warps/SM");
        run_hmm_loop<1, 4>(sms, 20, "PairHMM inner loop, two reads f32x2, K=4, 20 warps/SM");
        return 0;
    }
    if (!only_fp) run_all<0>(sms, d_out, d_cyc, bps);
    run_fp2<0>(sms, d_out, d_cyc, bps, "FFMA2 (packed f32x2)", 2, 0);
    run_fp2<1>(sms, d_out, d_cyc, bps, "FFMA2, scalar broadcast multiplier", 2, 0);
    run_fp2<2>(sms, d_out, d_cyc, bps, "mix 1 FFMA2 : 1 LOP3", 2, 1);
    run_fp2<5>(sms, d_out, d_cyc, bps, "mix 1 FFMA2 : 2 LOP3", 2, 2);
    run_fp2<3>(sms, d_out, d_cyc, bps, "mix 1 FFMA2 : 1 FFMA", 3, 0);
    run_fp2<4>(sms, d_out, d_cyc, bps, "mix 4 FFMA2 : 1 SHFL.UP", 2, 0.25);
    run_fp2_pattern<0, 8>(sms, d_out, d_cyc, 2, "PairHMM cell pattern f32x2 K=8, 3 packed instr/cell (16 warps/SM)");
    run_fp2_pattern<1, 8>(sms, d_out, d_cyc, 2, "PairHMM cell pattern f32x2 K=8, prior multiply unpacked (16 warps/SM)");
    run_fp2_pattern<0, 6>(sms, d_out, d_cyc, 2, "PairHMM cell pattern f32x2 K=6, 3 packed instr/cell (16 warps/SM)");
    run_fp2_pattern<0, 8>(sms, d_out, d_cyc, 1, "PairHMM cell pattern f32x2 K=8, 3 packed instr/cell (8 warps/SM)");
    run_fp_pattern<0, 8>(sms, d_out, d_cyc, 2, "FFMA, three distinct source registers (16 warps/SM)");
    run_fp_pattern<1, 8>(sms, d_out, d_cyc, 2, "PairHMM cell pattern K=8, 6 FP32 instr/cell (16 warps/SM)");
    run_fp_pattern<1, 6>(sms, d_out, d_cyc, 2, "PairHMM cell pattern K=6, 6 FP32 instr/cell (16 warps/SM)");
    run_fp_pattern<1, 8>(sms, d_out, d_cyc, 1, "PairHMM cell pattern K=8, 6 FP32 instr/cell (8 warps/SM)");
    return 0;
}
