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

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e = (x);                                                       \
        if (e != cudaSuccess) {                                                    \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

constexpr int NCHAIN = 8;
constexpr int UNROLL = 8;   // ops per chain per loop iteration

enum Op {
    OP_VIADDMNMX_S16X2, OP_VIMNMX3_S16X2_RELU, OP_VIADD_16X2, OP_VIMNMX_S16X2, OP_PRMT,
    OP_VIADDMNMX_S32, OP_VIMNMX3_S32_RELU, OP_IADD3, OP_LOP3, OP_IMAD,
    OP_FFMA, OP_FMUL, OP_FADD, OP_FSEL,
    OP_MIX_DPX_IMAD, OP_MIX_DPX_FFMA, OP_MIX_FFMA_SEL, OP_SWCELL, OP_SHFL,
    OP_MIX_VIADD_VIADDMNMX, OP_MIX_VIADDMNMX_VIMNMX3, OP_MIX_PRMT_VIADDMNMX, OP_MIX_VIADD_PRMT, OP_MIX_VIADD_IMAD,
    OP_MIX_VIMNMX3_IMAD, OP_MIX_PRMT_IMAD, OP_MIX_LOP3_VIADDMNMX, OP_MIX_VIADD_FFMA, OP_SWCELL_FULL, OP_COUNT
};
static const char *op_name[OP_COUNT] = {
    "VIADDMNMX.S16x2", "VIMNMX3.S16x2.RELU", "VIADD.16x2", "VIMNMX.S16x2", "PRMT",
    "VIADDMNMX.S32", "VIMNMX3.S32.RELU", "IADD3", "LOP3", "IMAD",
    "FFMA", "FMUL", "FADD", "FSEL(ISETP+SEL)",
    "mix 1 VIADDMNMX.S16x2 : 1 IMAD", "mix 1 VIADDMNMX.S16x2 : 1 FFMA", "mix 3 FFMA : 1 ISETP+FSEL",
    "SW s16x2 cell w/o PRMT (6 ops)", "SHFL.UP",
    "mix VIADD.16x2 : VIADDMNMX.S16x2", "mix VIADDMNMX.S16x2 : VIMNMX3.S16x2", "mix PRMT : VIADDMNMX.S16x2",
    "mix VIADD.16x2 : PRMT", "mix VIADD.16x2 : IMAD", "mix VIMNMX3.S16x2 : IMAD", "mix PRMT : IMAD",
    "mix LOP3 : VIADDMNMX.S16x2", "mix VIADD.16x2 : FFMA", "SW s16x2 cell with PRMT (7 ops)"};
// lane-ops counted per chain step
static const double op_count[OP_COUNT] = {1, 2, 1, 2, 1, 1, 2, 1, 1, 1, 1, 1, 1, 2, 2, 2, 5, 6, 1,
                                              2, 2, 2, 2, 2, 2, 2, 2, 2, 7};

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
    const double per_clk_sm = ops_per_thread * threads * blocks_per_sm / best_cyc;
    printf("{\"op\": \"%s\", \"lane_ops_per_clk_per_sm\": %.2f, \"tera_lane_ops_per_s\": %.3f, "
           "\"ms\": %.4f, \"eff_clock_mhz\": %.0f, \"warps_per_sm\": %d}\n",
           op_name[OP], per_clk_sm, lane_ops / (best_ms * 1e-3) / 1e12, best_ms,
           best_cyc / (best_ms * 1e-3) / 1e6, blocks_per_sm * threads / 32);
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
    run_all<0>(sms, d_out, d_cyc, bps);
    run_fp_pattern<0, 8>(sms, d_out, d_cyc, 2, "FFMA, three distinct source registers (16 warps/SM)");
    run_fp_pattern<1, 8>(sms, d_out, d_cyc, 2, "PairHMM cell pattern K=8, 6 FP32 instr/cell (16 warps/SM)");
    run_fp_pattern<1, 6>(sms, d_out, d_cyc, 2, "PairHMM cell pattern K=6, 6 FP32 instr/cell (16 warps/SM)");
    run_fp_pattern<1, 8>(sms, d_out, d_cyc, 1, "PairHMM cell pattern K=8, 6 FP32 instr/cell (8 warps/SM)");
    return 0;
}
