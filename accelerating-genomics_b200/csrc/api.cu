// api.cu -- C ABI of libagx.so (see include/agx.h): device contexts, the multi-GPU dispatcher
// (pair / read sharding by cell count, one host thread + one stream per GPU, no collective) and the
// host <-> device staging for the flat and pointer-array entry points.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <deque>
#include <vector>

#include "common.cuh"

namespace agx {

std::atomic<int64_t> g_launches{0};
std::atomic<int> g_profiling{0};

void ProfSpan::begin(cudaStream_t st)
{
    if (!g_profiling.load(std::memory_order_relaxed)) return;
    if (!e0) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
    cudaEventRecord(e0, st);
    armed = false;
}
void ProfSpan::end(cudaStream_t st)
{
    if (!g_profiling.load(std::memory_order_relaxed) || !e0) return;
    cudaEventRecord(e1, st);
    armed = true;
}
double ProfSpan::ms()
{
    if (!armed) return -1.0;
    if (cudaEventSynchronize(e1) != cudaSuccess) return -1.0;
    float v = 0.f;
    if (cudaEventElapsedTime(&v, e0, e1) != cudaSuccess) return -1.0;
    return (double)v;
}
void ProfSpan::destroy()
{
    if (e0) { cudaEventDestroy(e0); cudaEventDestroy(e1); }
    e0 = e1 = nullptr;
    armed = false;
}

static thread_local std::string t_error;
void set_error(const std::string &msg) { t_error = msg; }
int fail(int code, const std::string &msg)
{
    t_error = msg;
    return code;
}

namespace {

// Device address p with p + x == base + (x - bias): callers keep ABSOLUTE offsets (of a host buffer / file image)
// while only the byte range [bias, bias + size) is resident.  Formed with integer arithmetic -- pointer arithmetic
// may not leave the allocation -- and every access through it lands inside the buffer.
inline const uint8_t *biased(const void *base, int64_t bias)
{
    return reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(base) - (uintptr_t)bias);
}

// grow-only device buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return AGX_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return fail(AGX_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        cap = want;
        return AGX_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return AGX_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) return fail(AGX_ENOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
        cap = want;
        return AGX_OK;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

// stream + scratch + staging buffers of one in-flight alignment chunk (sw_align_shard)
struct AlignLane {
    cudaStream_t st = nullptr;         // the context's stream / its second lane's stream
    SwAlignWorkspace ws;
    DevBuf bytes, off, len, scores, ends, coords, cigar;
    int64_t pend_q0 = 0, pend_m = 0, pend_base = 0;   // chunk whose results are on their way home; its first run's index
};

// stream + workspace + staging buffers of one in-flight part of a PairHMM shard (hmm_shard)
struct HmmLane {
    cudaStream_t st = nullptr;         // the context's stream / its second lane's stream
    HmmWorkspace *ws = nullptr;        // the context's hmm / hmm_b
    DevBuf bytes, idx, out;
    PinBuf h_idx;                      // shard-local index arrays, built on the host
    PinBuf h_out;                      // results on their way to a PAGEABLE result array (a D2H copy into one would
    int64_t pend_o0 = 0, pend_n = 0;   // block the host until the kernels are done); pend_*: what still sits there
};

// stream + workspace + staging buffers of one in-flight SW chunk
struct SwLane {
    cudaStream_t st = nullptr;
    SwWorkspace ws;
    DevBuf bytes, off, len, out;
    PinBuf h_out;              // results land here first: a D2H copy into pageable user memory would block
    int64_t pend_q0 = 0, pend_m = 0;   // chunk whose results still sit in h_out
};

struct DeviceCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    SwParseWorkspace parse[2];
    cudaStream_t copy_stream = nullptr;    // uploads of sw_score_file_image
    cudaStream_t prep_stream = nullptr;    // high priority: chunking + length classes of the next region
    cudaEvent_t lane_done[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> seg_events;   // upload segments of sw_score_file_image
    SwLane lane[2];          // lane[0].st == stream
    SwWorkspace &sw = lane[0].ws;
    HmmWorkspace hmm, hmm_b;                 // hmm_b / hmm_parse_b: second lane of pairhmm_forward_file_image
    HmmParseWorkspace hmm_parse, hmm_parse_b;
    HmmLane hl[2];                           // pairhmm_forward_batches_flat: two parts of a shard in flight
    DevBuf d_bytes, d_a, d_b, d_c, d_d, d_e, d_f, d_out;   // staging for the host entry points
    AlignLane al[2];                                        // sw_ends_* / sw_align_*: two chunks in flight
    PinBuf h_cigar;                                         // several GPUs: the CIGAR runs of this GPU's shard
    SwAlignWorkspace &align = al[0].ws;                     // (the device-resident entry points use the first)
    int64_t align_free = 0;                                 // free device memory + what the alignment scratch holds, as last asked
    PinBuf h_a, h_b, h_out;
    std::string error;   // error raised on this device's worker thread
    std::string name;    // agx_device_name()
};

std::mutex g_mu;
std::vector<std::unique_ptr<DeviceCtx>> g_ctx;
std::atomic<int> g_gatk{0}, g_force64{0};

DeviceCtx *ctx_for_device(int device)
{
    for (auto &c : g_ctx)
        if (c->device == device) return c.get();
    return nullptr;
}

int init_devices(const std::vector<int> &ids)
{
    std::lock_guard<std::mutex> lk(g_mu);
    int n_visible = 0;
    cudaError_t e = cudaGetDeviceCount(&n_visible);
    if (e != cudaSuccess || n_visible == 0)
        return fail(AGX_ENODEVICE, std::string("no CUDA device: ") +
                                       (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    std::vector<int> want = ids;
    if (want.empty())
        for (int i = 0; i < n_visible; ++i) want.push_back(i);
    for (int d : want)
        if (d < 0 || d >= n_visible) return fail(AGX_EINVAL, "agx_init: device ordinal out of range");
    for (int d : want) {
        if (ctx_for_device(d)) continue;
        cudaDeviceProp prop;
        AGX_CUDA(cudaGetDeviceProperties(&prop, d));
        if (prop.major < 10)
            return fail(AGX_ENODEVICE, std::string("device ") + prop.name +
                                           " is not sm_100-class; libagx carries sm_100a code only");
        auto c = std::make_unique<DeviceCtx>();
        c->device = d;
        AGX_CUDA(cudaSetDevice(d));
        AGX_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->lane[0].st = c->stream;
        AGX_CUDA(cudaStreamCreateWithFlags(&c->lane[1].st, cudaStreamNonBlocking));
        AGX_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        int prio_lo = 0, prio_hi = 0;
        AGX_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        AGX_CUDA(cudaStreamCreateWithPriority(&c->prep_stream, cudaStreamNonBlocking, prio_hi));
        for (cudaEvent_t &ev : c->lane_done) AGX_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        g_ctx.push_back(std::move(c));
    }
    return AGX_OK;
}

// Split [0, n) into `parts` contiguous ranges of roughly equal weight.
std::vector<int64_t> balanced_cuts(const std::vector<double> &prefix, int parts)
{
    const int64_t n = (int64_t)prefix.size() - 1;
    std::vector<int64_t> cuts(parts + 1, n);
    cuts[0] = 0;
    const double total = prefix[n];
    for (int k = 1; k < parts; ++k) {
        const double target = total * k / parts;
        cuts[k] = std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin();
        if (cuts[k] < cuts[k - 1]) cuts[k] = cuts[k - 1];
        if (cuts[k] > n) cuts[k] = n;
    }
    return cuts;
}

// The same for pairs weighted by their cell count, without a prefix array of the whole batch (8 MB of fresh pages
// per 10^6 pairs): sums over blocks of 1024 pairs, then a scan inside the block a cut falls into.
std::vector<int64_t> balanced_pair_cuts(const int32_t *len, int64_t n_pairs, int parts)
{
    const int64_t B = 1024, nb = (n_pairs + B - 1) / B;
    auto weight = [&](int64_t p) { return (double)(len[2 * p] + 1) * (double)(len[2 * p + 1] + 1); };
    std::vector<double> block(nb + 1, 0.0);
    for (int64_t b = 0; b < nb; ++b) {
        double w = 0.0;
        for (int64_t p = b * B, e = std::min(n_pairs, p + B); p < e; ++p) w += weight(p);
        block[b + 1] = block[b] + w;
    }
    std::vector<int64_t> cuts(parts + 1, n_pairs);
    cuts[0] = 0;
    const double total = block[nb];
    for (int k = 1; k < parts; ++k) {
        const double target = total * k / parts;
        int64_t b = std::lower_bound(block.begin(), block.end(), target) - block.begin();     // block[b] >= target
        b = std::max<int64_t>(b - 1, 0);
        double w = block[b];
        int64_t p = b * B;
        while (p < n_pairs && w < target) w += weight(p++);
        cuts[k] = std::max(std::min(p, n_pairs), cuts[k - 1]);
    }
    return cuts;
}

// Run fn(ctx, shard_index) on every configured device, one host thread each.
template <typename Fn> int for_each_device(int n_shards, Fn fn)
{
    std::vector<int> rcs(n_shards, AGX_OK);
    std::vector<std::string> errs(n_shards);
    auto body = [&](int k) {
        DeviceCtx *c = g_ctx[k].get();
        if (cudaSetDevice(c->device) != cudaSuccess) {
            rcs[k] = AGX_ECUDA;
            errs[k] = "cudaSetDevice failed";
            return;
        }
        rcs[k] = fn(*c, k);
        if (rcs[k] != AGX_OK) errs[k] = t_error;
    };
    if (n_shards == 1) {
        body(0);
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < n_shards; ++k) th.emplace_back(body, k);
        for (auto &t : th) t.join();
    }
    for (int k = 0; k < n_shards; ++k)
        if (rcs[k] != AGX_OK) return fail(rcs[k], "gpu " + std::to_string(g_ctx[k]->device) + ": " + errs[k]);
    return AGX_OK;
}

int require_init()
{
    if (g_ctx.empty()) {
        int rc = init_devices({});
        if (rc != AGX_OK) return rc;
    }
    return AGX_OK;
}

// Validates every (offset, length) of a flat batch before any device work and returns the longest sequence.
// Branch-free passes over slices of the arrays, on up to four host threads for large batches (2 * 10^6 entries take
// 1.3 ms on one core -- as long as a fifth of the batch takes the GPU).
int validate_sequences(const char *what, int64_t seqs_bytes, const int64_t *off, const int32_t *len, int64_t n_seqs,
                       int32_t *longest_out)
{
    struct Acc { int32_t longest = 0, shortest = 0; int64_t min_off = 0, max_end = 0; };
    const int parts = n_seqs >= ((int64_t)1 << 19) ? 4 : 1;
    Acc acc[4];
    auto pass = [&](int k) {
        Acc a;
        for (int64_t i = n_seqs * k / parts, e = n_seqs * (k + 1) / parts; i < e; ++i) {
            a.longest = std::max(a.longest, len[i]);
            a.shortest = std::min(a.shortest, len[i]);
            a.min_off = std::min(a.min_off, off[i]);
            a.max_end = std::max(a.max_end, off[i] + len[i]);
        }
        acc[k] = a;
    };
    if (parts == 1) {
        pass(0);
    } else {
        std::thread th[3];
        for (int k = 1; k < parts; ++k) th[k - 1] = std::thread(pass, k);
        pass(0);
        for (int k = 1; k < parts; ++k) th[k - 1].join();
    }
    Acc t = acc[0];
    for (int k = 1; k < parts; ++k) {
        t.longest = std::max(t.longest, acc[k].longest);
        t.shortest = std::min(t.shortest, acc[k].shortest);
        t.min_off = std::min(t.min_off, acc[k].min_off);
        t.max_end = std::max(t.max_end, acc[k].max_end);
    }
    if (t.shortest < 0 || t.min_off < 0 || t.max_end > seqs_bytes)
        for (int64_t i = 0; i < n_seqs; ++i)
            if (len[i] < 0 || off[i] < 0 || off[i] + len[i] > seqs_bytes)
                return fail(AGX_EINVAL, std::string(what) + ": sequence " + std::to_string(i) + " lies outside the buffer");
    *longest_out = t.longest;
    return AGX_OK;
}

// ---------------------------------------------------------------- SW on one shard
// One shard = a contiguous range of pairs on one GPU.  It is cut into chunks that alternate between the
// context's two lanes (stream + workspace + staging buffers each).  The host->device copies of chunk k+1 are
// queued BEFORE the host blocks on chunk k's grid-sizing read-back, so the link never waits for the host and the
// DP kernels run under the copies; results go straight to a pinned result array (a pageable one is filled from a
// pinned staging buffer).  Offsets are uploaded as they are; the device base pointer is shifted by the chunk's
// first byte instead.
int sw_shard(DeviceCtx &c, const uint8_t *seqs, int64_t seqs_bytes, const int64_t *off, const int32_t *len,
             int64_t p0, int64_t p1, SwScoring sc, int32_t *scores_out)
{
    const int64_t n = p1 - p0;
    if (n <= 0) return AGX_OK;
    // chunks: an eighth of the shard, 64 Ki .. 512 Ki pairs (what is left to do when the last byte has landed is one
    // chunk's kernels); one chunk up to 96 Ki pairs
    int64_t chunk = n;
    if (n > 98304) {
        chunk = (n + 7) / 8;
        if (chunk < 65536) chunk = 65536;
        if (chunk > 524288) chunk = 524288;
    }
    if (const char *e = getenv("AGX_SW_CHUNK")) {        // tuning knob: pairs per chunk (0 = one chunk)
        const long long v = atoll(e);
        chunk = v > 0 ? std::min<int64_t>(v, n) : n;
    }
    const int64_t n_chunks = (n + chunk - 1) / chunk;
    // byte range [lo, hi) every chunk's sequences lie in (offsets were validated by sw_flat_impl); the passes over
    // the offsets run on up to four host threads
    std::vector<int64_t> lo(n_chunks, 0), hi(n_chunks, 0);
    {
        auto range_of = [&](int64_t k) {
            const int64_t q0 = p0 + k * chunk, q1 = std::min(p1, q0 + chunk);
            int64_t l = INT64_MAX, h = 0;
            for (int64_t i = 2 * q0; i < 2 * q1; ++i) {
                l = std::min(l, off[i]);
                h = std::max(h, off[i] + len[i]);
            }
            if (h < l) { l = 0; h = 0; }
            lo[k] = l;
            hi[k] = h;
        };
        const int parts = (n >= ((int64_t)1 << 18) && n_chunks > 1) ? (int)std::min<int64_t>(4, n_chunks) : 1;
        auto slice = [&](int t) { for (int64_t k = t; k < n_chunks; k += parts) range_of(k); };
        std::thread th[3];
        for (int t = 1; t < parts; ++t) th[t - 1] = std::thread(slice, t);
        slice(0);
        for (int t = 1; t < parts; ++t) th[t - 1].join();
    }
    cudaPointerAttributes attr;
    const bool out_pinned = cudaPointerGetAttributes(&attr, scores_out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();

    auto stage = [&](int64_t k) -> int {
        SwLane &L = c.lane[k & 1];
        const int64_t q0 = p0 + k * chunk, q1 = std::min(p1, q0 + chunk), m = q1 - q0;
        const int64_t l = lo[k], h = hi[k];
        int rc;
        if ((rc = L.bytes.reserve((size_t)(h - l) + 16)) != AGX_OK) return rc;
        if ((rc = L.off.reserve((size_t)m * 2 * sizeof(int64_t))) != AGX_OK) return rc;
        if ((rc = L.len.reserve((size_t)m * 2 * sizeof(int32_t))) != AGX_OK) return rc;
        if ((rc = L.out.reserve((size_t)m * sizeof(int32_t))) != AGX_OK) return rc;
        AGX_CUDA(cudaMemcpyAsync(L.bytes.p, seqs + l, (size_t)(h - l), cudaMemcpyHostToDevice, L.st));
        AGX_CUDA(cudaMemcpyAsync(L.off.p, off + 2 * q0, (size_t)m * 2 * sizeof(int64_t), cudaMemcpyHostToDevice, L.st));
        AGX_CUDA(cudaMemcpyAsync(L.len.p, len + 2 * q0, (size_t)m * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, L.st));
        return AGX_OK;
    };

    // results of a lane's previous chunk: wait for them and (pageable result array) hand them to the caller
    auto drain = [&](SwLane &L) -> int {
        if (L.pend_m == 0) return AGX_OK;
        AGX_CUDA(cudaStreamSynchronize(L.st));
        if (!out_pinned) memcpy(scores_out + L.pend_q0, L.h_out.p, (size_t)L.pend_m * sizeof(int32_t));
        L.pend_m = 0;
        return AGX_OK;
    };
    c.lane[0].pend_m = c.lane[1].pend_m = 0;

    int rc = stage(0);
    if (rc != AGX_OK) return rc;
    for (int64_t k = 0; k < n_chunks; ++k) {
        SwLane &L = c.lane[k & 1];
        const int64_t q0 = p0 + k * chunk, q1 = std::min(p1, q0 + chunk), m = q1 - q0;
        if (k + 1 < n_chunks) {
            // the other lane still owns chunk k-1: collect its results (its kernels ran under chunk k's upload), then
            // queue chunk k+1's upload behind chunk k's
            if ((rc = drain(c.lane[(k + 1) & 1])) != AGX_OK) return rc;
            if ((rc = stage(k + 1)) != AGX_OK) return rc;
        }
        if (!out_pinned && (rc = L.h_out.reserve((size_t)m * sizeof(int32_t))) != AGX_OK) return rc;
        // blocks until chunk k is on the device and classified, then queues its DP kernels
        rc = sw_run_device(L.ws, biased(L.bytes.p, lo[k]), L.off.as<int64_t>(), L.len.as<int32_t>(), m, sc,
                           L.out.as<int32_t>(), L.st);
        if (rc != AGX_OK) return rc;
        AGX_CUDA(cudaMemcpyAsync(out_pinned ? (void *)(scores_out + q0) : L.h_out.p, L.out.p, (size_t)m * sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, L.st));
        L.pend_q0 = q0;
        L.pend_m = m;
    }
    if ((rc = drain(c.lane[0])) != AGX_OK) return rc;
    if ((rc = drain(c.lane[1])) != AGX_OK) return rc;
    AGX_CUDA(cudaStreamSynchronize(c.lane[0].st));
    AGX_CUDA(cudaStreamSynchronize(c.lane[1].st));
    return AGX_OK;
}

int sw_flat_impl(const uint8_t *seqs, int64_t seqs_bytes, const int64_t *off, const int32_t *len,
                 int64_t n_pairs, SwScoring sc, int32_t *scores_out)
{
    if (n_pairs < 0) return fail(AGX_EINVAL, "sw: n_pairs < 0");
    if (n_pairs == 0) return AGX_OK;
    if (!seqs || !off || !len || !scores_out) return fail(AGX_EINVAL, "sw: null argument");
    int32_t longest = 0;
    int rc = validate_sequences("sw", seqs_bytes, off, len, 2 * n_pairs, &longest);
    if (rc != AGX_OK) return rc;
    rc = require_init();
    if (rc != AGX_OK) return rc;

    // very long alignments: one at a time, columns striped over every configured GPU (sw_long.cu)
    std::vector<int64_t> giants;
    const bool maybe_giant = (int64_t)longest * (int64_t)longest >= sw_long_cells() && longest > 1025;
    for (int64_t p = 0; maybe_giant && p < n_pairs; ++p)
        if ((int64_t)len[2 * p] * (int64_t)len[2 * p + 1] >= sw_long_cells() &&
            std::min(len[2 * p], len[2 * p + 1]) > 1025)
            giants.push_back(p);
    if (!giants.empty()) {
        if (!(sc.match > 0 && sc.mismatch < 0 && sc.gap_open <= 0 && sc.gap_extend < 0))
            return fail(AGX_ERANGE, "sw: scoring must satisfy match > 0 > mismatch, gap_open <= 0, gap_extend < 0");
        std::vector<int> devs;
        std::vector<cudaStream_t> sts;
        std::vector<SwLongWorkspace *> wss;
        for (auto &c : g_ctx) { devs.push_back(c->device); sts.push_back(c->stream); wss.push_back(&c->sw.lng); }
        for (int64_t p : giants) {
            // columns = the longer sequence (more stripes in flight); AGX_LONG_SWAP flips that for experiments
            int hi = len[2 * p] >= len[2 * p + 1] ? 0 : 1;
            if (getenv("AGX_LONG_SWAP")) hi = 1 - hi;
            rc = sw_long_host_multi((int)devs.size(), devs.data(), sts.data(), wss.data(), seqs + off[2 * p + hi],
                                    len[2 * p + hi], seqs + off[2 * p + 1 - hi], len[2 * p + 1 - hi], sc,
                                    scores_out + p);
            if (rc != AGX_OK) return rc;
        }
        if ((int64_t)giants.size() == n_pairs) return AGX_OK;
        // score the remaining pairs as an ordinary batch
        std::vector<int64_t> off2;
        std::vector<int32_t> len2;
        std::vector<int64_t> idx;
        size_t gi = 0;
        for (int64_t p = 0; p < n_pairs; ++p) {
            if (gi < giants.size() && giants[gi] == p) { ++gi; continue; }
            idx.push_back(p);
            off2.push_back(off[2 * p]); off2.push_back(off[2 * p + 1]);
            len2.push_back(len[2 * p]); len2.push_back(len[2 * p + 1]);
        }
        std::vector<int32_t> sc2(idx.size());
        rc = sw_flat_impl(seqs, seqs_bytes, off2.data(), len2.data(), (int64_t)idx.size(), sc, sc2.data());
        if (rc != AGX_OK) return rc;
        for (size_t k = 0; k < idx.size(); ++k) scores_out[idx[k]] = sc2[k];
        return AGX_OK;
    }

    const int n_dev = (int)std::min<int64_t>((int64_t)g_ctx.size(), n_pairs);
    std::vector<int64_t> cuts{0, n_pairs};
    if (n_dev > 1) cuts = balanced_pair_cuts(len, n_pairs, n_dev);
    return for_each_device(n_dev, [&](DeviceCtx &c, int k) {
        return sw_shard(c, seqs, seqs_bytes, off, len, cuts[k], cuts[k + 1], sc, scores_out);
    });
}

// Device memory the traceback matrices may count on: what is free + what the alignment scratch already holds.
// cudaMemGetInfo() is not cheap (measured up to milliseconds next to a caching allocator), so the answer is kept
// and asked again only after a call ran out of memory.
int align_memory(DeviceCtx &c, bool refresh, int64_t *out)
{
    if (refresh || c.align_free == 0) {
        size_t fr = 0, tot = 0;
        AGX_CUDA(cudaMemGetInfo(&fr, &tot));
        c.align_free = (int64_t)fr;
        for (const AlignLane &L : c.al) c.align_free += L.ws.cap_tb + L.ws.cap_tb_gen;
    }
    *out = c.align_free;
    return AGX_OK;
}

// ---------------------------------------------------------------- SW alignment (end cell / start cell / CIGAR)
// One shard = a contiguous range of pairs on one GPU, cut into chunks whose H-byte matrices fit the budget.
// mode 1: scores + ends; mode 2: + coords and CIGAR runs (appended to `cigar`, run offsets relative to the shard).
struct AlignOut {
    int32_t *scores, *ends, *coords;
    int64_t *cigar_off;               // mode 2: [n] run offsets relative to the shard (+ the total as entry n when `whole`)
    bool whole;                       // the shard is the whole batch (one GPU)
    uint32_t *cigar_direct;           // mode 2, one GPU: the caller's array, runs land there as long as they fit
    int64_t cigar_cap;
    int64_t *total;                   // mode 2, several GPUs: runs of this shard; they wait in the context's pinned
};                                    // h_cigar for the caller, who knows their place once every shard is done

int sw_align_shard(DeviceCtx &c, const uint8_t *seqs, const int64_t *off, const int32_t *len, int64_t p0, int64_t p1,
                   SwScoring sc, int mode, AlignOut out)
{
    // Two lanes (stream + scratch + staging buffers each) alternate over the chunks: chunk k+1 is cut and its upload
    // queued BEFORE the host blocks on chunk k's kernels (the call returns the run total, so it waits for them), and
    // chunk k's results travel home while chunk k+1 computes.  The host waits for a lane only when it comes back to
    // it two chunks later.
    AlignLane *lanes[2] = {&c.al[0], &c.al[1]};
    lanes[0]->st = c.stream;
    lanes[1]->st = c.lane[1].st;
    const int64_t n = p1 - p0;
    int64_t budget = (int64_t)8 << 30;
    auto tb_budget = [&](bool refresh) -> int {
        if (mode != 2) return AGX_OK;
        // a quarter of (what is free + what the scratch already holds) per lane, at most 40 GiB
        // (a bound, not an allocation: the scratch grows to what the chunks really take)
        int64_t avail = 0;
        int rc0 = align_memory(c, refresh, &avail);
        if (rc0 != AGX_OK) return rc0;
        budget = std::min<int64_t>(avail / 4, (int64_t)40 << 30);
        if (const char *e = getenv("AGX_ALIGN_TB_BYTES")) budget = std::max<long long>(atoll(e), 1 << 20);
        return AGX_OK;
    };
    int rc = tb_budget(false);
    if (rc != AGX_OK) return rc;
    // pairs per chunk: a quarter of the shard, 64 Ki .. 256 Ki pairs (one chunk up to 64 Ki pairs); a short first
    // chunk, to start computing earlier, was measured and loses to its own fixed costs (profiles/r2bc_*)
    int64_t chunk_pairs = n <= 65536 ? n : std::min<int64_t>(262144, std::max<int64_t>(65536, (n + 3) / 4));
    if (const char *e = getenv("AGX_ALIGN_CHUNK")) {         // tuning knob: pairs per chunk (0 = as many as fit)
        const long long v = atoll(e);
        chunk_pairs = v > 0 ? v : (int64_t)1 << 20;
    }
    int shrink = 0;
    c.align.dp_ms_sum = c.align.walk_ms_sum = -1.0;
    const bool trace = getenv("AGX_ALIGN_TRACE") != nullptr;      // host-side timeline of the chunks on stderr
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();

    struct Chunk { int64_t q0, q1, l, h; };
    // chunk: at most chunk_pairs pairs and, in mode 2, matrices within the budget (the class layout rounds rows up to
    // the longest of the class, so the per-pair bound is scaled by what the chunk's longest line adds); the same
    // pass finds the byte range [l, h) the chunk's sequences lie in
    int32_t tb_cap = 0;
    const int32_t *tb_tab = sw_align_tb_row_table(&tb_cap);
    auto row_bytes_of = [&](int32_t la, int32_t lb) -> int64_t {
        const int32_t ra = la > lb ? lb : la;
        return ra <= tb_cap ? tb_tab[ra] : (int64_t)((ra + 255) / 256) * 256;
    };
    auto cut = [&](int64_t q0) {
        Chunk ch{q0, q0, INT64_MAX, 0};
        const int64_t most = std::max<int64_t>(1, chunk_pairs >> shrink);
        const int64_t qe = std::min(p1, q0 + most);
        // the usual case in one branch-free pass: all `most` pairs fit
        int64_t rows = 0, l = INT64_MAX, h = 0;
        int32_t lng = 1;
        for (int64_t q = q0; q < qe; ++q) {
            const int32_t la = len[2 * q], lb = len[2 * q + 1];
            const int64_t oa = off[2 * q], ob = off[2 * q + 1];
            if (mode == 2) rows += row_bytes_of(la, lb);
            lng = std::max(lng, std::max(la, lb));
            l = std::min(l, std::min(oa, ob));
            h = std::max(h, std::max(oa + la, ob + lb));
        }
        if (mode != 2 || qe - q0 <= 1 || (double)rows * (lng + 32) * 1.02 <= (double)(budget >> shrink)) {
            ch.q1 = qe; ch.l = l; ch.h = h;
        } else {
            int64_t row_bytes = 0;
            int32_t longest_c = 1;
            while (ch.q1 < qe) {
                const int32_t la = len[2 * ch.q1], lb = len[2 * ch.q1 + 1];
                const int32_t new_longest = std::max(longest_c, std::max(la, lb));
                const int64_t grown = row_bytes + row_bytes_of(la, lb);
                // every pair of a class is padded to the class's longest row sequence (+ the systolic skew)
                if (ch.q1 > q0 && (double)grown * (new_longest + 32) * 1.02 > (double)(budget >> shrink)) break;
                row_bytes = grown;
                longest_c = new_longest;
                const int64_t oa = off[2 * ch.q1], ob = off[2 * ch.q1 + 1];
                ch.l = std::min(ch.l, std::min(oa, ob));
                ch.h = std::max(ch.h, std::max(oa + la, ob + lb));
                ++ch.q1;
            }
        }
        if (ch.h < ch.l) { ch.l = 0; ch.h = 0; }
        return ch;
    };
    // results of a lane's previous chunk: wait for them; run offsets of later chunks are rebased on the host
    int64_t cig_held = 0;                             // runs already on their way into h_cigar
    auto drain = [&](AlignLane &L) -> int {
        if (L.pend_m == 0) return AGX_OK;
        AGX_CUDA(cudaStreamSynchronize(L.st));
        if (mode == 2 && L.pend_base != 0)
            for (int64_t i = 0; i < L.pend_m; ++i) out.cigar_off[(L.pend_q0 - p0) + i] += L.pend_base;
        if (g_profiling.load()) {
            const double d = L.ws.prof_dp.ms(), w = mode == 2 ? L.ws.prof_walk.ms() : -1.0;
            if (d >= 0) c.align.dp_ms_sum = std::max(c.align.dp_ms_sum, 0.0) + d;
            if (w >= 0) c.align.walk_ms_sum = std::max(c.align.walk_ms_sum, 0.0) + w;
        }
        if (trace) fprintf(stderr, "[agx align] +%.2f ms: results of pairs [%lld, %lld) on the host\n", now() - t_begin,
                           (long long)L.pend_q0, (long long)(L.pend_q0 + L.pend_m));
        L.pend_m = 0;
        return AGX_OK;
    };
    // (no host wait: the uploads queue up behind the lane's previous results on their way home)
    auto stage = [&](AlignLane &L, const Chunk &ch) -> int {
        int r;
        const int64_t m = ch.q1 - ch.q0;
        if ((r = L.bytes.reserve((size_t)(ch.h - ch.l) + 16)) != AGX_OK) return r;
        if ((r = L.off.reserve((size_t)m * 2 * sizeof(int64_t))) != AGX_OK) return r;
        if ((r = L.len.reserve((size_t)m * 2 * sizeof(int32_t))) != AGX_OK) return r;
        if ((r = L.scores.reserve((size_t)m * sizeof(int32_t))) != AGX_OK) return r;
        if ((r = L.ends.reserve((size_t)m * 2 * sizeof(int32_t))) != AGX_OK) return r;
        if (mode == 2 && (r = L.coords.reserve((size_t)m * 4 * sizeof(int32_t))) != AGX_OK) return r;
        // only the bytes [l, h) are resident; the offsets go up as they are and the kernels address a biased base
        AGX_CUDA(cudaMemcpyAsync(L.bytes.p, seqs + ch.l, (size_t)(ch.h - ch.l), cudaMemcpyHostToDevice, L.st));
        AGX_CUDA(cudaMemcpyAsync(L.off.p, off + 2 * ch.q0, (size_t)m * 2 * sizeof(int64_t), cudaMemcpyHostToDevice, L.st));
        AGX_CUDA(cudaMemcpyAsync(L.len.p, len + 2 * ch.q0, (size_t)m * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, L.st));
        if (trace) fprintf(stderr, "[agx align] +%.2f ms: pairs [%lld, %lld) cut, upload queued\n", now() - t_begin,
                           (long long)ch.q0, (long long)ch.q1);
        return AGX_OK;
    };

    // several GPUs: the runs collect in pinned memory of the context (grown with its contents kept; both lanes idle then)
    auto keep_runs = [&](int64_t runs) -> int {
        const size_t need = (size_t)runs * sizeof(uint32_t);
        if (need <= c.h_cigar.cap) return AGX_OK;
        AGX_CUDA(cudaStreamSynchronize(lanes[0]->st));
        AGX_CUDA(cudaStreamSynchronize(lanes[1]->st));
        PinBuf bigger;
        int r = bigger.reserve(std::max(need * 2, (size_t)16 << 20));
        if (r != AGX_OK) return r;
        if (c.h_cigar.p && cig_held > 0) memcpy(bigger.p, c.h_cigar.p, (size_t)cig_held * sizeof(uint32_t));
        c.h_cigar.release();
        c.h_cigar = bigger;
        return AGX_OK;
    };
    int64_t cig_base = 0;
    int li = 0;
    c.al[0].pend_m = c.al[1].pend_m = 0;
    Chunk cur = cut(p0), pre{p1, p1, 0, 0};
    bool pre_ok = false;
    if ((rc = stage(*lanes[li], cur)) != AGX_OK) return rc;
    while (cur.q0 < p1) {
        AlignLane &L = *lanes[li];
        const int64_t m = cur.q1 - cur.q0;
        if ((rc = drain(L)) != AGX_OK) return rc;      // the chunk before last (long home by now): offsets, profiling spans
        Chunk nxt{p1, p1, 0, 0};
        if (cur.q1 < p1) {
            nxt = (pre_ok && pre.q0 == cur.q1) ? pre : cut(cur.q1);
            if ((rc = stage(*lanes[li ^ 1], nxt)) != AGX_OK) return rc;
        }
        // the chunk after next is cut by a helper thread while this one computes (the host is about to block)
        pre_ok = false;
        std::thread helper;
        if (nxt.q1 < p1) helper = std::thread([&cut, &pre, q = nxt.q1] { pre = cut(q); });
        int64_t total = 0;
        rc = sw_align_run_device(L.ws, biased(L.bytes.p, cur.l), L.off.as<int64_t>(), L.len.as<int32_t>(), m, sc, mode, budget,
                                 L.scores.as<int32_t>(), L.ends.as<int32_t>(), L.coords.as<int32_t>(), &total, L.st);
        if (helper.joinable()) { helper.join(); pre_ok = rc == AGX_OK; }     // (cut with this shrink / budget)
        if (rc == AGX_ENOMEM && mode == 2 && m > 1 && shrink < 24) {     // the bound was too optimistic (or memory went elsewhere)
            if ((rc = tb_budget(true)) != AGX_OK) return rc;
            ++shrink;
            cur = cut(cur.q0);               // (the chunk staged on the other lane is cut again after this one)
            if ((rc = stage(L, cur)) != AGX_OK) return rc;
            continue;
        }
        if (rc != AGX_OK) return rc;
        if (trace) fprintf(stderr, "[agx align] +%.2f ms: kernels of pairs [%lld, %lld) done (%lld runs)\n", now() - t_begin,
                           (long long)cur.q0, (long long)cur.q1, (long long)total);
        AGX_CUDA(cudaMemcpyAsync(out.scores + cur.q0, L.scores.p, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, L.st));
        if (out.ends)
            AGX_CUDA(cudaMemcpyAsync(out.ends + 2 * cur.q0, L.ends.p, (size_t)m * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, L.st));
        if (mode == 2) {
            AGX_CUDA(cudaMemcpyAsync(out.coords + 4 * cur.q0, L.coords.p, (size_t)m * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, L.st));
            // (m entries: the entry after them belongs to the next chunk, or is the shard's total set below)
            AGX_CUDA(cudaMemcpyAsync(out.cigar_off + (cur.q0 - p0), L.ws.cig_off, (size_t)m * sizeof(int64_t),
                                     cudaMemcpyDeviceToHost, L.st));
            const bool fits = !out.cigar_direct || cig_base + total <= out.cigar_cap;
            if (total > 0 && fits) {
                if ((rc = L.cigar.reserve((size_t)total * sizeof(uint32_t))) != AGX_OK) return rc;
                if ((rc = sw_align_gather_device(L.ws, m, L.cigar.as<uint32_t>(), L.st)) != AGX_OK) return rc;
                uint32_t *dst = out.cigar_direct ? out.cigar_direct + cig_base : nullptr;
                if (!dst) {
                    if ((rc = keep_runs(cig_base + total)) != AGX_OK) return rc;
                    dst = c.h_cigar.as<uint32_t>() + cig_base;
                    cig_held = cig_base + total;
                }
                AGX_CUDA(cudaMemcpyAsync(dst, L.cigar.p, (size_t)total * sizeof(uint32_t), cudaMemcpyDeviceToHost, L.st));
            }
        }
        L.pend_q0 = cur.q0;
        L.pend_m = m;
        L.pend_base = cig_base;
        cig_base += total;
        cur = nxt;
        li ^= 1;
    }
    if ((rc = drain(c.al[0])) != AGX_OK) return rc;
    if ((rc = drain(c.al[1])) != AGX_OK) return rc;
    if (mode == 2) {
        if (out.whole) out.cigar_off[n] = cig_base;
        if (out.total) *out.total = cig_base;
    }
    if (trace) fprintf(stderr, "[agx align] +%.2f ms: shard done\n", now() - t_begin);
    return AGX_OK;
}

int sw_align_impl(const uint8_t *seqs, int64_t seqs_bytes, const int64_t *off, const int32_t *len, int64_t n_pairs,
                  SwScoring sc, int mode, int32_t *scores_out, int32_t *ends_out, int32_t *coords_out,
                  int64_t *cigar_off_out, uint32_t *cigar_out, int64_t cigar_cap, int64_t *cigar_total_out)
{
    if (cigar_total_out) *cigar_total_out = 0;
    if (n_pairs < 0) return fail(AGX_EINVAL, "sw align: n_pairs < 0");
    if (n_pairs == 0) { if (mode == 2 && cigar_off_out) cigar_off_out[0] = 0; return AGX_OK; }
    if (!seqs || !off || !len || !scores_out) return fail(AGX_EINVAL, "sw align: null argument");
    if (mode == 1 && !ends_out) return fail(AGX_EINVAL, "sw ends: null argument");
    if (mode == 2 && (!coords_out || !cigar_off_out || (!cigar_out && cigar_cap > 0) || cigar_cap < 0))
        return fail(AGX_EINVAL, "sw align: null argument");
    int32_t longest = 0;
    int rc = validate_sequences("sw align", seqs_bytes, off, len, 2 * n_pairs, &longest);
    if (rc != AGX_OK) return rc;
    rc = require_init();
    if (rc != AGX_OK) return rc;
    if (mode == 1) {
        // whole-GPU pairs: score and end cell from the striped long-alignment kernel, one pair at a time over every
        // configured GPU (the traceback of such a pair would need its 10^12-cell matrix: mode 2 refuses them)
        std::vector<int64_t> giants;
        const int64_t long_cells = sw_long_cells();
        const bool maybe_giant = (int64_t)longest * (int64_t)longest >= long_cells && longest > 1025;
        for (int64_t p = 0; maybe_giant && p < n_pairs; ++p)
            if ((int64_t)len[2 * p] * (int64_t)len[2 * p + 1] >= long_cells && std::min(len[2 * p], len[2 * p + 1]) > 1025)
                giants.push_back(p);
        if (!giants.empty()) {
            if (!(sc.match > 0 && sc.mismatch < 0 && sc.gap_open <= 0 && sc.gap_extend < 0))
                return fail(AGX_ERANGE, "sw: scoring must satisfy match > 0 > mismatch, gap_open <= 0, gap_extend < 0");
            std::vector<int> devs;
            std::vector<cudaStream_t> sts;
            std::vector<SwLongWorkspace *> wss;
            for (auto &c : g_ctx) { devs.push_back(c->device); sts.push_back(c->stream); wss.push_back(&c->sw.lng); }
            for (int64_t p : giants) {
                const int hi = len[2 * p] >= len[2 * p + 1] ? 0 : 1;          // columns = the longer sequence
                const int sx = len[2 * p] > len[2 * p + 1] ? 1 : 0;           // antidiagonalSmithWaterman.c:229
                int32_t ec = -1, er = -1;
                rc = sw_long_host_multi((int)devs.size(), devs.data(), sts.data(), wss.data(), seqs + off[2 * p + hi],
                                        len[2 * p + hi], seqs + off[2 * p + 1 - hi], len[2 * p + 1 - hi], sc, scores_out + p,
                                        &ec, &er, sx == hi);
                if (rc != AGX_OK) return rc;
                ends_out[2 * p + hi] = ec;
                ends_out[2 * p + 1 - hi] = er;
            }
            if ((int64_t)giants.size() == n_pairs) return AGX_OK;
            std::vector<int64_t> off2, idx;
            std::vector<int32_t> len2;
            size_t gi = 0;
            for (int64_t p = 0; p < n_pairs; ++p) {
                if (gi < giants.size() && giants[gi] == p) { ++gi; continue; }
                idx.push_back(p);
                off2.push_back(off[2 * p]); off2.push_back(off[2 * p + 1]);
                len2.push_back(len[2 * p]); len2.push_back(len[2 * p + 1]);
            }
            std::vector<int32_t> sc2(idx.size()), en2(2 * idx.size());
            rc = sw_align_impl(seqs, seqs_bytes, off2.data(), len2.data(), (int64_t)idx.size(), sc, 1, sc2.data(), en2.data(),
                               nullptr, nullptr, nullptr, 0, nullptr);
            if (rc != AGX_OK) return rc;
            for (size_t k = 0; k < idx.size(); ++k) {
                scores_out[idx[k]] = sc2[k];
                ends_out[2 * idx[k]] = en2[2 * k];
                ends_out[2 * idx[k] + 1] = en2[2 * k + 1];
            }
            return AGX_OK;
        }
    }
    const int n_dev = (int)std::min<int64_t>((int64_t)g_ctx.size(), n_pairs);
    std::vector<int64_t> cuts{0, n_pairs};
    if (n_dev > 1) cuts = balanced_pair_cuts(len, n_pairs, n_dev);
    std::vector<int64_t> totals(n_dev, 0);
    rc = for_each_device(n_dev, [&](DeviceCtx &c, int k) {
        AlignOut o;
        o.scores = scores_out;
        o.ends = mode == 2 ? nullptr : ends_out;       // mode 2: the ends are part of coords
        o.coords = coords_out;
        o.cigar_off = mode == 2 ? cigar_off_out + cuts[k] : nullptr;
        o.whole = n_dev == 1;
        o.cigar_direct = n_dev == 1 ? cigar_out : nullptr;
        o.cigar_cap = n_dev == 1 ? cigar_cap : INT64_MAX;
        o.total = &totals[k];
        return sw_align_shard(c, seqs, off, len, cuts[k], cuts[k + 1], sc, mode, o);
    });
    if (rc != AGX_OK) return rc;
    if (mode != 2) return AGX_OK;
    int64_t base = totals[0];
    if (n_dev > 1) {
        // every shard's runs wait in its context's pinned memory and its offsets count from its own first run: each
        // GPU's host thread moves them to their place in the caller's arrays
        std::vector<int64_t> bases(n_dev + 1, 0);
        for (int k = 0; k < n_dev; ++k) bases[k + 1] = bases[k] + totals[k];
        base = bases[n_dev];
        rc = for_each_device(n_dev, [&](DeviceCtx &c, int k) {
            if (bases[k] != 0)
                for (int64_t i = cuts[k]; i < cuts[k + 1]; ++i) cigar_off_out[i] += bases[k];
            if (totals[k] > 0 && bases[k + 1] <= cigar_cap)
                memcpy(cigar_out + bases[k], c.h_cigar.p, (size_t)totals[k] * sizeof(uint32_t));
            return (int)AGX_OK;
        });
        if (rc != AGX_OK) return rc;
    }
    cigar_off_out[n_pairs] = base;
    if (cigar_total_out) *cigar_total_out = base;
    if (base > cigar_cap)
        return fail(AGX_ERANGE, "sw align: " + std::to_string(base) + " CIGAR runs do not fit cigar_cap = " +
                                    std::to_string(cigar_cap) + " (scores, coordinates and offsets are complete)");
    return AGX_OK;
}

// ---------------------------------------------------------------- PairHMM on one shard
struct HmmHost {
    const uint8_t *buf;
    int64_t buf_bytes;
    const int64_t *read_field_off;
    const int32_t *read_len;
    int64_t n_reads;
    const int64_t *hap_off;
    const int32_t *hap_len;
    int64_t n_haps;
    const int64_t *batch_read_start;
    const int64_t *batch_hap_start;
    int64_t n_batches;
    std::vector<int32_t> read_batch;   // [n_reads]
    std::vector<int64_t> read_out_off; // [n_reads]
    int64_t n_pairs;
};

// Waits for what a lane has in flight and hands staged results to a pageable result array.
int hmm_lane_wait(HmmLane &L, double *out)
{
    const cudaError_t e = cudaStreamSynchronize(L.st);
    if (e == cudaSuccess && L.pend_n > 0) memcpy(out + L.pend_o0, L.h_out.p, (size_t)L.pend_n * sizeof(double));
    L.pend_n = 0;
    AGX_CUDA(e);
    return AGX_OK;
}

// Reads [r0, r1) on one lane (stream + workspace + staging buffers).  wait = true: the call returns with the results
// in `out`; false: everything is queued on the lane's stream (the FP64 rescue reads its work count on the device, so
// no host round trip follows the stream kernels) and the caller synchronises the lane before it uses it again.
int hmm_part(HmmLane &L, const HmmHost &h, int64_t r0, int64_t r1, double *out, bool wait)
{
    const int64_t nr = r1 - r0;
    if (nr <= 0) return AGX_OK;
    const int32_t b0 = h.read_batch[r0], b1 = h.read_batch[r1 - 1] + 1;
    const int64_t hh0 = h.batch_hap_start[b0], hh1 = h.batch_hap_start[b1];
    const int64_t nh = hh1 - hh0, nb = b1 - b0;
    // outputs of this shard are contiguous: [o0, o1)
    const int64_t o0 = h.read_out_off[r0];
    const int64_t last_b = h.read_batch[r1 - 1];
    const int64_t o1 = h.read_out_off[r1 - 1] + (h.batch_hap_start[last_b + 1] - h.batch_hap_start[last_b]);
    const int64_t n_out = o1 - o0;
    if (n_out <= 0) return AGX_OK;

    int64_t lo = INT64_MAX, hi = 0;
    for (int64_t r = r0; r < r1; ++r)
        for (int f = 0; f < 5; ++f) {
            lo = std::min(lo, h.read_field_off[5 * r + f]);
            hi = std::max(hi, h.read_field_off[5 * r + f] + h.read_len[r]);
        }
    for (int64_t x = hh0; x < hh1; ++x) {
        lo = std::min(lo, h.hap_off[x]);
        hi = std::max(hi, h.hap_off[x] + h.hap_len[x]);
    }
    const int64_t nbytes = hi - lo;

    // shard-local index arrays, packed into one pinned block
    const size_t sz_rfo = (size_t)nr * 5 * sizeof(int64_t), sz_roo = (size_t)nr * sizeof(int64_t),
                 sz_ho = (size_t)nh * sizeof(int64_t), sz_bhs = (size_t)(nb + 1) * sizeof(int64_t),
                 sz_rl = (size_t)nr * sizeof(int32_t), sz_rb = (size_t)nr * sizeof(int32_t),
                 sz_hl = (size_t)nh * sizeof(int32_t);
    const size_t total = sz_rfo + sz_roo + sz_ho + sz_bhs + sz_rl + sz_rb + sz_hl;
    int rc;
    if ((rc = L.h_idx.reserve(total)) != AGX_OK) return rc;
    if ((rc = L.idx.reserve(total)) != AGX_OK) return rc;
    if ((rc = L.bytes.reserve((size_t)nbytes + 16)) != AGX_OK) return rc;
    if ((rc = L.out.reserve((size_t)n_out * sizeof(double))) != AGX_OK) return rc;
    uint8_t *hp = L.h_idx.as<uint8_t>();
    int64_t *p_rfo = (int64_t *)hp;
    int64_t *p_roo = (int64_t *)(hp + sz_rfo);
    int64_t *p_ho = (int64_t *)(hp + sz_rfo + sz_roo);
    int64_t *p_bhs = (int64_t *)(hp + sz_rfo + sz_roo + sz_ho);
    int32_t *p_rl = (int32_t *)(hp + sz_rfo + sz_roo + sz_ho + sz_bhs);
    int32_t *p_rb = (int32_t *)(hp + sz_rfo + sz_roo + sz_ho + sz_bhs + sz_rl);
    int32_t *p_hl = (int32_t *)(hp + sz_rfo + sz_roo + sz_ho + sz_bhs + sz_rl + sz_rb);
    for (int64_t r = 0; r < nr; ++r) {
        for (int f = 0; f < 5; ++f) p_rfo[5 * r + f] = h.read_field_off[5 * (r0 + r) + f] - lo;
        p_roo[r] = h.read_out_off[r0 + r] - o0;
        p_rl[r] = h.read_len[r0 + r];
        p_rb[r] = h.read_batch[r0 + r] - b0;
    }
    for (int64_t x = 0; x < nh; ++x) {
        p_ho[x] = h.hap_off[hh0 + x] - lo;
        p_hl[x] = h.hap_len[hh0 + x];
    }
    for (int64_t b = 0; b <= nb; ++b) p_bhs[b] = h.batch_hap_start[b0 + b] - hh0;

    cudaStream_t st = L.st;
    AGX_CUDA(cudaMemcpyAsync(L.bytes.p, h.buf + lo, (size_t)nbytes, cudaMemcpyHostToDevice, st));
    AGX_CUDA(cudaMemcpyAsync(L.idx.p, hp, total, cudaMemcpyHostToDevice, st));
    uint8_t *dp = L.idx.as<uint8_t>();
    HmmBatchView v;
    v.buf = L.bytes.as<uint8_t>();
    v.read_field_off = (const int64_t *)dp;
    const int64_t *d_roo = (const int64_t *)(dp + sz_rfo);
    v.hap_off = (const int64_t *)(dp + sz_rfo + sz_roo);
    v.batch_hap_start = (const int64_t *)(dp + sz_rfo + sz_roo + sz_ho);
    v.read_len = (const int32_t *)(dp + sz_rfo + sz_roo + sz_ho + sz_bhs);
    v.read_batch = (const int32_t *)(dp + sz_rfo + sz_roo + sz_ho + sz_bhs + sz_rl);
    v.hap_len = (const int32_t *)(dp + sz_rfo + sz_roo + sz_ho + sz_bhs + sz_rl + sz_rb);
    v.n_reads = nr;
    v.n_haps = nh;
    v.n_batches = nb;
    rc = hmm_run_device(*L.ws, v, nbytes, d_roo, n_out, g_gatk.load(), g_force64.load() != 0, wait ? 1 : 2,
                        L.out.as<double>(), st);
    if (rc != AGX_OK) return rc;
    cudaPointerAttributes attr;
    const bool out_pinned = cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (out_pinned) {
        AGX_CUDA(cudaMemcpyAsync(out + o0, L.out.p, (size_t)n_out * sizeof(double), cudaMemcpyDeviceToHost, st));
    } else {
        if ((rc = L.h_out.reserve((size_t)n_out * sizeof(double))) != AGX_OK) return rc;
        AGX_CUDA(cudaMemcpyAsync(L.h_out.p, L.out.p, (size_t)n_out * sizeof(double), cudaMemcpyDeviceToHost, st));
        L.pend_o0 = o0;
        L.pend_n = n_out;
    }
    if (wait) return hmm_lane_wait(L, out);
    return AGX_OK;
}

// ---------------------------------------------------------------- PairHMM on one shard
// One shard = a contiguous range of reads on one GPU.  A large shard is cut into parts that alternate between two
// lanes: while the stream kernels of part k run, the host builds the index arrays of part k+1 and its upload goes
// out, and the results of part k travel home under part k+1 (the call used to upload, compute and download in turn).
int hmm_shard(DeviceCtx &c, const HmmHost &h, int64_t r0, int64_t r1, double *out)
{
    const int64_t nr = r1 - r0;
    if (nr <= 0) return AGX_OK;
    c.hl[0].st = c.stream;
    c.hl[0].ws = &c.hmm;
    c.hl[1].st = c.lane[1].st;
    c.hl[1].ws = &c.hmm_b;
    // parts of at least 32 Ki reads (a part must still fill the GPU: ~10^4 two-read warps over five row classes), four at
    // most (200 000 reads: 28.2 ms in one part, 25.8 / 24.4 / 24.7 / 26.1 ms in 2 / 4 / 6 / 8, profiles/r2bl_hmm_flat.jsonl)
    int parts = (int)std::max<int64_t>(1, std::min<int64_t>(4, nr / 32768));
    if (const char *e = getenv("AGX_HMM_PARTS")) parts = std::max(1, atoi(e));       // tuning knob
    parts = (int)std::min<int64_t>(parts, nr);
    if (parts == 1) return hmm_part(c.hl[0], h, r0, r1, out, true);
    int rc = AGX_OK;
    for (int k = 0; k < parts && rc == AGX_OK; ++k) {
        HmmLane &L = c.hl[k & 1];
        // the lane's previous part (two parts back) must be done before its buffers are filled again; the part in
        // between keeps the GPU busy meanwhile
        if (k >= 2 && (rc = hmm_lane_wait(L, out)) != AGX_OK) break;
        rc = hmm_part(L, h, r0 + nr * k / parts, r0 + nr * (k + 1) / parts, out, false);
    }
    const int w0 = hmm_lane_wait(c.hl[0], out), w1 = hmm_lane_wait(c.hl[1], out);
    if (rc != AGX_OK) return rc;
    return w0 != AGX_OK ? w0 : w1;
}


int hmm_flat_impl(HmmHost &h, double *out)
{
    if (h.n_batches < 0 || h.n_reads < 0 || h.n_haps < 0) return fail(AGX_EINVAL, "pairhmm: negative count");
    if (h.n_batches == 0 || h.n_reads == 0) return AGX_OK;
    if (!h.buf || !h.read_field_off || !h.read_len || !h.hap_off || !h.hap_len || !h.batch_read_start ||
        !h.batch_hap_start || !out)
        return fail(AGX_EINVAL, "pairhmm: null argument");
    if (h.batch_read_start[0] != 0 || h.batch_hap_start[0] != 0 ||
        h.batch_read_start[h.n_batches] != h.n_reads || h.batch_hap_start[h.n_batches] != h.n_haps)
        return fail(AGX_EINVAL, "pairhmm: batch_*_start must run from 0 to n_reads / n_haps");
    for (int64_t r = 0; r < h.n_reads; ++r) {
        if (h.read_len[r] < 1 || h.read_len[r] > 8192)
            return fail(AGX_ERANGE, "pairhmm: read " + std::to_string(r) + " length outside [1, 8192]");
        for (int f = 0; f < 5; ++f) {
            const int64_t o = h.read_field_off[5 * r + f];
            if (o < 0 || o + h.read_len[r] > h.buf_bytes)
                return fail(AGX_EINVAL, "pairhmm: read " + std::to_string(r) + " lies outside the buffer");
        }
    }
    for (int64_t x = 0; x < h.n_haps; ++x) {
        if (h.hap_len[x] < 1 || h.hap_len[x] > (1 << 20))
            return fail(AGX_ERANGE, "pairhmm: haplotype " + std::to_string(x) + " length outside [1, 2^20]");
        if (h.hap_off[x] < 0 || h.hap_off[x] + h.hap_len[x] > h.buf_bytes)
            return fail(AGX_EINVAL, "pairhmm: haplotype " + std::to_string(x) + " lies outside the buffer");
    }
    int rc = require_init();
    if (rc != AGX_OK) return rc;

    h.read_batch.resize(h.n_reads);
    h.read_out_off.resize(h.n_reads);
    std::vector<double> prefix(h.n_reads + 1, 0.0);
    int64_t o = 0;
    for (int64_t b = 0; b < h.n_batches; ++b) {
        const int64_t rs = h.batch_read_start[b], re = h.batch_read_start[b + 1];
        const int64_t hs = h.batch_hap_start[b], he = h.batch_hap_start[b + 1];
        if (re < rs || he < hs) return fail(AGX_EINVAL, "pairhmm: batch_*_start must be non-decreasing");
        double hap_cells = 0;
        for (int64_t x = hs; x < he; ++x) hap_cells += h.hap_len[x];
        for (int64_t r = rs; r < re; ++r) {
            h.read_batch[r] = (int32_t)b;
            h.read_out_off[r] = o;
            o += he - hs;
            prefix[r + 1] = prefix[r] + hap_cells * h.read_len[r];
        }
    }
    h.n_pairs = o;
    if (o == 0) return AGX_OK;
    const int n_dev = (int)std::min<int64_t>((int64_t)g_ctx.size(), h.n_reads);
    std::vector<int64_t> cuts{0, h.n_reads};
    if (n_dev > 1) cuts = balanced_cuts(prefix, n_dev);
    return for_each_device(n_dev, [&](DeviceCtx &c, int k) {
        return hmm_shard(c, h, cuts[k], cuts[k + 1], out);
    });
}

// One host thread per shard, each on its own GPU and stream; returns the first failure.
template <typename Shard, typename Fn> int for_each_shard(const Shard *shards, int32_t n_shards, Fn fn)
{
    if (n_shards < 0 || (n_shards > 0 && !shards)) return fail(AGX_EINVAL, "shards: null argument");
    std::vector<DeviceCtx *> ctx(n_shards, nullptr);
    for (int k = 0; k < n_shards; ++k) {
        ctx[k] = ctx_for_device(shards[k].device);
        if (!ctx[k]) return fail(AGX_ENODEVICE, "shards: device " + std::to_string(shards[k].device) + " was not passed to agx_init");
        for (int j = 0; j < k; ++j)
            if (ctx[j] == ctx[k]) return fail(AGX_EINVAL, "shards: device " + std::to_string(shards[k].device) + " appears twice");
    }
    std::vector<int> rcs(n_shards, AGX_OK);
    std::vector<std::string> errs(n_shards);
    auto body = [&](int k) {
        if (cudaSetDevice(ctx[k]->device) != cudaSuccess) { rcs[k] = AGX_ECUDA; errs[k] = "cudaSetDevice failed"; return; }
        rcs[k] = fn(*ctx[k], shards[k]);
        if (rcs[k] == AGX_OK && cudaStreamSynchronize(ctx[k]->stream) != cudaSuccess) { rcs[k] = AGX_ECUDA; t_error = "stream synchronisation failed"; }
        if (rcs[k] != AGX_OK) errs[k] = t_error;
    };
    if (n_shards == 1) {
        body(0);
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < n_shards; ++k) th.emplace_back(body, k);
        for (auto &t : th) t.join();
    }
    for (int k = 0; k < n_shards; ++k)
        if (rcs[k] != AGX_OK) return fail(rcs[k], "gpu " + std::to_string(shards[k].device) + ": " + errs[k]);
    return AGX_OK;
}

}  // namespace
}  // namespace agx

using namespace agx;

// =============================================================================== C ABI
extern "C" {

int agx_init(int32_t n_gpus)
{
    std::vector<int> ids;
    if (n_gpus > 0)
        for (int i = 0; i < n_gpus; ++i) ids.push_back(i);
    return init_devices(ids);
}

int agx_init_devices(const int32_t *device_ids, int32_t n_devices)
{
    if (n_devices <= 0 || !device_ids) return fail(AGX_EINVAL, "agx_init_devices: empty device list");
    return init_devices(std::vector<int>(device_ids, device_ids + n_devices));
}

int32_t agx_device_count(void) { return (int32_t)g_ctx.size(); }

int32_t agx_device_ordinal(int32_t index)
{
    return (index >= 0 && index < (int32_t)g_ctx.size()) ? g_ctx[index]->device : -1;
}

const char *agx_device_name(int32_t index)
{
    if (index < 0 || index >= (int32_t)g_ctx.size()) return "";
    DeviceCtx &c = *g_ctx[index];
    if (c.name.empty()) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, c.device) == cudaSuccess) c.name = prop.name;
    }
    return c.name.c_str();
}

void agx_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto &c : g_ctx) {
        cudaSetDevice(c->device);
        if (c->stream) cudaStreamSynchronize(c->stream);
        if (c->lane[1].st) cudaStreamSynchronize(c->lane[1].st);
        for (SwLane &L : c->lane) {
            sw_workspace_free(L.ws);
            for (DevBuf *b : {&L.bytes, &L.off, &L.len, &L.out}) b->release();
            L.h_out.release();
        }
        if (c->lane[1].st) cudaStreamDestroy(c->lane[1].st);
        sw_parse_workspace_free(c->parse[0]);
        sw_parse_workspace_free(c->parse[1]);
        if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
        if (c->prep_stream) cudaStreamDestroy(c->prep_stream);
        for (cudaEvent_t ev : c->lane_done) if (ev) cudaEventDestroy(ev);
        for (cudaEvent_t ev : c->seg_events) cudaEventDestroy(ev);
        hmm_workspace_free(c->hmm);
        hmm_workspace_free(c->hmm_b);
        hmm_parse_workspace_free(c->hmm_parse);
        hmm_parse_workspace_free(c->hmm_parse_b);
        for (DevBuf *b : {&c->d_bytes, &c->d_a, &c->d_b, &c->d_c, &c->d_d, &c->d_e, &c->d_f, &c->d_out}) b->release();
        for (PinBuf *b : {&c->h_a, &c->h_b, &c->h_out, &c->h_cigar}) b->release();
        for (HmmLane &L : c->hl) {
            for (DevBuf *b : {&L.bytes, &L.idx, &L.out}) b->release();
            L.h_idx.release();
            L.h_out.release();
        }
        for (AlignLane &L : c->al) {
            for (DevBuf *b : {&L.bytes, &L.off, &L.len, &L.scores, &L.ends, &L.coords, &L.cigar}) b->release();
            sw_align_workspace_free(L.ws);
        }
        if (c->stream) cudaStreamDestroy(c->stream);
    }
    g_ctx.clear();
}

const char *agx_last_error(void) { return t_error.c_str(); }
const char *agx_version(void) { return "agx 0.1 sm_100a"; }
int64_t agx_launch_count(void) { return g_launches.load(); }
void agx_reset_launch_count(void) { g_launches.store(0); }

int agx_set_profiling(int32_t on) { g_profiling.store(on ? 1 : 0); return AGX_OK; }

double agx_profile_ms(int32_t device, int32_t which)
{
    DeviceCtx *c = ctx_for_device(device);
    if (!c) return -1.0;
    cudaSetDevice(device);
    switch (which) {
    case AGX_PROF_SW_DUO: return c->sw.prof_duo.ms();
    case AGX_PROF_SW_WAVE: return c->sw.prof_wave.ms();
    case AGX_PROF_SW_CLASSIFY: return c->sw.prof_classify.ms();
    case AGX_PROF_HMM_STREAM: return c->hmm.prof_stream.ms();
    case AGX_PROF_HMM_FP64: return c->hmm.prof_fp64.ms();
    case AGX_PROF_HMM_CLASSIFY: return c->hmm.prof_classify.ms();
    case AGX_PROF_SW_ALIGN_DP: return c->align.dp_ms_sum;
    case AGX_PROF_SW_ALIGN_WALK: return c->align.walk_ms_sum;
    case AGX_PROF_SW_LONG: { const double v = c->sw.lng.prof.ms(); return v >= 0 ? v : c->sw.prof_long.ms(); }
    default: return -1.0;
    }
}

int agx_pairhmm_set_gatk_mode(int32_t on) { g_gatk.store(on & 3); return AGX_OK; }
int agx_pairhmm_set_force_fp64(int32_t on) { g_force64.store(on ? 1 : 0); return AGX_OK; }

int sw_score_batch_flat(const uint8_t *seqs, int64_t seqs_bytes, const int64_t *off, const int32_t *len,
                        int64_t n_pairs, int32_t match, int32_t mismatch, int32_t gap_open,
                        int32_t gap_extend, int32_t *scores_out)
{
    return sw_flat_impl(seqs, seqs_bytes, off, len, n_pairs, SwScoring{match, mismatch, gap_open, gap_extend},
                        scores_out);
}

int sw_ends_batch_flat(const uint8_t *seqs, int64_t seqs_bytes, const int64_t *off, const int32_t *len, int64_t n_pairs,
                       int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend, int32_t *scores_out,
                       int32_t *ends_out)
{
    return sw_align_impl(seqs, seqs_bytes, off, len, n_pairs, SwScoring{match, mismatch, gap_open, gap_extend}, 1,
                         scores_out, ends_out, nullptr, nullptr, nullptr, 0, nullptr);
}

int sw_align_batch_flat(const uint8_t *seqs, int64_t seqs_bytes, const int64_t *off, const int32_t *len, int64_t n_pairs,
                        int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend, int32_t *scores_out,
                        int32_t *coords_out, int64_t *cigar_off_out, uint32_t *cigar_out, int64_t cigar_cap,
                        int64_t *cigar_total_out)
{
    return sw_align_impl(seqs, seqs_bytes, off, len, n_pairs, SwScoring{match, mismatch, gap_open, gap_extend}, 2,
                         scores_out, nullptr, coords_out, cigar_off_out, cigar_out, cigar_cap, cigar_total_out);
}

// pointer-array forms: gathered into one flat image (a0 b0 a1 b1 ...), then the flat entry points
static int gather_pairs(const uint8_t *const *a, const int32_t *a_len, const uint8_t *const *b, const int32_t *b_len,
                        int64_t n_pairs, std::vector<uint8_t> &flat, std::vector<int64_t> &off, std::vector<int32_t> &len)
{
    if (n_pairs < 0) return fail(AGX_EINVAL, "sw: n_pairs < 0");
    if (n_pairs > 0 && (!a || !a_len || !b || !b_len)) return fail(AGX_EINVAL, "sw: null argument");
    off.resize((size_t)n_pairs * 2);
    len.resize((size_t)n_pairs * 2);
    int64_t total = 0;
    for (int64_t p = 0; p < n_pairs; ++p) {
        if (a_len[p] < 0 || b_len[p] < 0 || (a_len[p] && !a[p]) || (b_len[p] && !b[p]))
            return fail(AGX_EINVAL, "sw: bad sequence " + std::to_string(p));
        off[2 * p] = total; len[2 * p] = a_len[p];
        total += a_len[p];
        off[2 * p + 1] = total; len[2 * p + 1] = b_len[p];
        total += b_len[p];
    }
    flat.resize((size_t)total + 1);
    for (int64_t p = 0; p < n_pairs; ++p) {
        if (a_len[p]) memcpy(flat.data() + off[2 * p], a[p], (size_t)a_len[p]);
        if (b_len[p]) memcpy(flat.data() + off[2 * p + 1], b[p], (size_t)b_len[p]);
    }
    return AGX_OK;
}

int sw_ends_batch(const uint8_t *const *a, const int32_t *a_len, const uint8_t *const *b, const int32_t *b_len,
                  int64_t n_pairs, int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                  int32_t *scores_out, int32_t *ends_out)
{
    std::vector<uint8_t> flat;
    std::vector<int64_t> off;
    std::vector<int32_t> len;
    int rc = gather_pairs(a, a_len, b, b_len, n_pairs, flat, off, len);
    if (rc != AGX_OK) return rc;
    return sw_ends_batch_flat(flat.data(), (int64_t)flat.size() - 1, off.data(), len.data(), n_pairs, match, mismatch,
                              gap_open, gap_extend, scores_out, ends_out);
}

int sw_align_batch(const uint8_t *const *a, const int32_t *a_len, const uint8_t *const *b, const int32_t *b_len,
                   int64_t n_pairs, int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                   int32_t *scores_out, int32_t *coords_out, int64_t *cigar_off_out, uint32_t *cigar_out,
                   int64_t cigar_cap, int64_t *cigar_total_out)
{
    std::vector<uint8_t> flat;
    std::vector<int64_t> off;
    std::vector<int32_t> len;
    int rc = gather_pairs(a, a_len, b, b_len, n_pairs, flat, off, len);
    if (rc != AGX_OK) return rc;
    return sw_align_batch_flat(flat.data(), (int64_t)flat.size() - 1, off.data(), len.data(), n_pairs, match, mismatch,
                               gap_open, gap_extend, scores_out, coords_out, cigar_off_out, cigar_out, cigar_cap,
                               cigar_total_out);
}

// sw_score_file_image over several GPUs: the image is cut into one byte range per GPU at line starts; every
// GPU uploads and chunks its own range (phase 1, one host thread per GPU), the host turns the chunk counts
// into the global pairing -- a range that starts at an odd chunk index scores its first chunk against the
// last chunk of the range before it, whose bytes it uploaded as well -- and every GPU scores the pairs whose
// second chunk it holds (phase 2).  Returns 1 when the input does not suit this path (the caller then takes
// the single-GPU one).
static int sw_file_image_multi(const uint8_t *image, int64_t image_bytes, int64_t hlen, int32_t line_buf, SwScoring sc,
                        int64_t want_pairs, int32_t *scores_out, int64_t scores_cap, int64_t *n_pairs_out,
                        int64_t *dangling_off, int32_t *dangling_len)
{
    const int n_dev = (int)g_ctx.size();
    const int64_t body = image_bytes - hlen;
    if (n_dev < 2 || body < (int64_t)n_dev * (4 << 20)) return 1;
    std::vector<int64_t> cut(n_dev + 1, image_bytes);
    cut[0] = hlen;
    for (int d = 1; d < n_dev; ++d) {
        const int64_t nominal = hlen + body * d / n_dev;
        const void *nl = memchr(image + nominal, '\n', (size_t)(image_bytes - nominal));
        cut[d] = nl ? (const uint8_t *)nl - image + 1 : image_bytes;
        if (cut[d] < cut[d - 1]) cut[d] = cut[d - 1];
    }
    struct Part { int64_t n_chunks = 0, last_off = -1; int32_t last_len = 0; int64_t *d_off = nullptr; int32_t *d_len = nullptr; };
    std::vector<Part> part(n_dev);
    const int64_t max_chunks = 2 * want_pairs;
    int rc = for_each_device(n_dev, [&](DeviceCtx &c, int d) -> int {
        SwLane &L = c.lane[0];
        int r;
        if (cut[d + 1] <= cut[d]) return AGX_OK;
        // this GPU holds its own byte range only (+ the chunk before it); offsets stay those of the whole image
        // (from a multiple of 16: the newline scan reads the image in aligned 128-bit words)
        const int64_t up0 = std::max<int64_t>(hlen, cut[d] - (line_buf - 1)) & ~(int64_t)15;
        if ((r = L.bytes.reserve((size_t)(cut[d + 1] - up0) + 64)) != AGX_OK) return r;
        AGX_CUDA(cudaMemcpyAsync(L.bytes.p, image + up0, (size_t)(cut[d + 1] - up0), cudaMemcpyHostToDevice, L.st));
        return sw_parse_device(c.parse[0], biased(L.bytes.p, up0), cut[d], cut[d + 1], line_buf, max_chunks,
                               image[cut[d + 1] - 1], &part[d].d_off, &part[d].d_len, &part[d].n_chunks, &part[d].last_off,
                               &part[d].last_len, L.st);
    });
    if (rc != AGX_OK) return rc;
    for (int d = 0; d < n_dev; ++d)
        if (part[d].n_chunks == 0) return 1;            // e.g. one line longer than a range: not for this path
    // global chunk numbering, truncated to what line 1 asks for
    std::vector<int64_t> start(n_dev + 1, 0), used(n_dev, 0);
    for (int d = 0; d < n_dev; ++d) {
        used[d] = std::max<int64_t>(0, std::min(part[d].n_chunks, max_chunks - start[d]));
        start[d + 1] = start[d] + used[d];
    }
    const int64_t total = start[n_dev], n_pairs = total / 2;
    if (n_pairs > scores_cap)
        return fail(AGX_ERANGE, "sw: scores_out holds " + std::to_string(scores_cap) + " scores, the file has " +
                                    std::to_string(n_pairs) + " pairs");
    rc = for_each_device(n_dev, [&](DeviceCtx &c, int d) -> int {
        SwLane &L = c.lane[0];
        const int64_t p_lo = start[d] / 2, p_hi = (start[d] + used[d]) / 2, m = p_hi - p_lo;
        if (m <= 0) return AGX_OK;
        const int shift = (int)(start[d] & 1);          // odd: pair p_lo = (last chunk of the range before, my chunk 0)
        if (shift) {
            AGX_CUDA(cudaMemcpyAsync(part[d].d_off - 1, &part[d - 1].last_off, sizeof(int64_t), cudaMemcpyHostToDevice, L.st));
            AGX_CUDA(cudaMemcpyAsync(part[d].d_len - 1, &part[d - 1].last_len, sizeof(int32_t), cudaMemcpyHostToDevice, L.st));
        }
        int r;
        if ((r = L.out.reserve((size_t)m * sizeof(int32_t))) != AGX_OK) return r;
        if ((r = L.h_out.reserve((size_t)m * sizeof(int32_t))) != AGX_OK) return r;
        const int64_t up0 = std::max<int64_t>(hlen, cut[d] - (line_buf - 1)) & ~(int64_t)15;
        r = sw_run_device(L.ws, biased(L.bytes.p, up0), part[d].d_off - shift, part[d].d_len - shift, m, sc, L.out.as<int32_t>(), L.st);
        if (r != AGX_OK) return r;
        AGX_CUDA(cudaMemcpyAsync(L.h_out.p, L.out.p, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, L.st));
        AGX_CUDA(cudaStreamSynchronize(L.st));
        memcpy(scores_out + p_lo, L.h_out.p, (size_t)m * sizeof(int32_t));
        return AGX_OK;
    });
    if (rc != AGX_OK) return rc;
    if (want_pairs > n_pairs && (total & 1)) {
        // EOF in the middle of a pair: the dangling first line is the last chunk of the file
        if (dangling_off) *dangling_off = part[n_dev - 1].last_off;
        if (dangling_len) *dangling_len = part[n_dev - 1].last_len;
    }
    if (n_pairs_out) *n_pairs_out = n_pairs;
    return AGX_OK;
}

// The image goes to the device in segments on a copy stream; line-aligned regions are chunked (sw_parse.cu)
// and scored on the compute stream as soon as the segment that completes them has landed, so the DP
// kernels run under the remaining host->device copies.  A region that ends in the middle of a pair hands
// its last chunk to the next region (fgets() keeps no state besides the file position, so chunking may
// restart at any chunk start).
int sw_score_file_image(const uint8_t *image, int64_t image_bytes, int32_t line_buf, int32_t match,
                        int32_t mismatch, int32_t gap_open, int32_t gap_extend, int32_t *scores_out,
                        int64_t scores_cap, int64_t *n_pairs_out, int32_t *header_out, int64_t *dangling_off,
                        int32_t *dangling_len)
{
    if (n_pairs_out) *n_pairs_out = 0;
    if (header_out) *header_out = 0;
    if (dangling_off) *dangling_off = -1;
    if (dangling_len) *dangling_len = 0;
    if (!image || image_bytes <= 0) return fail(AGX_EINVAL, "sw: empty file image");
    if (line_buf < 2) return fail(AGX_EINVAL, "sw: line buffer must hold at least one character");
    // the header is the first fgets() chunk (antidiagonalSmithWaterman.c:205-209)
    const int64_t cap = line_buf - 1;
    const int64_t lim = std::min<int64_t>(image_bytes, cap);
    const void *nl = memchr(image, '\n', (size_t)lim);
    const int64_t hlen = nl ? (const uint8_t *)nl - image + 1 : lim;
    char tmp[32];
    const size_t c = (size_t)std::min<int64_t>(hlen, (int64_t)sizeof tmp - 1);
    memcpy(tmp, image, c);
    tmp[c] = 0;
    const int line_num = atoi(tmp);
    if (header_out) *header_out = line_num;
    const int64_t want_pairs = line_num > 0 ? ((int64_t)line_num + 1) / 2 : 0;
    if (want_pairs == 0 || hlen >= image_bytes) return AGX_OK;
    if (!scores_out) return fail(AGX_EINVAL, "sw: null argument");
    int rc = require_init();
    if (rc != AGX_OK) return rc;
    const SwScoring sc_multi{match, mismatch, gap_open, gap_extend};
    rc = sw_file_image_multi(image, image_bytes, hlen, line_buf, sc_multi, want_pairs, scores_out, scores_cap, n_pairs_out,
                             dangling_off, dangling_len);
    if (rc != 1) return rc;                  // done (or failed) on several GPUs; 1 = take the single-GPU path
    DeviceCtx &ctx = *g_ctx[0];
    AGX_CUDA(cudaSetDevice(ctx.device));
    SwLane &L = ctx.lane[0];                 // owns the image and the scores on the device
    cudaStream_t copy_st = ctx.copy_stream;
    const SwScoring sc{match, mismatch, gap_open, gap_extend};

    // every chunk holds at least one byte: an upper bound on the pairs this image can yield
    const int64_t max_pairs = std::min<int64_t>(want_pairs, (image_bytes - hlen) / 2 + 1);
    if ((rc = L.bytes.reserve((size_t)image_bytes + 64)) != AGX_OK) return rc;
    if ((rc = L.out.reserve((size_t)max_pairs * sizeof(int32_t))) != AGX_OK) return rc;

    // ---- upload in segments, one event per segment ------------------------------------------------
    // Every region costs the host three stream round trips (newline count, chunk count, length-class
    // counts: ~0.3 ms in all), so regions must be few and none shorter than the ~16 MiB that arrive in
    // that time; what is left to do when the last byte lands is one region, so they must not be long either:
    // about eleven equal segments, 16 .. 32 MiB.  Segment boundaries are multiples of 1 MiB: a copy that starts
    // at an odd byte runs visibly slower (image / 8 = an odd size cost 0.35 ms per 302 MB, profiles/r2ay_*).
    std::vector<int64_t> seg_end;
    {
        const int64_t mib = (int64_t)1 << 20;
        int64_t uniform = std::max(16 * mib, std::min(32 * mib, image_bytes / 11 / mib * mib));
        if (const char *e = getenv("AGX_SW_IMAGE_SEGMENT")) uniform = std::max<int64_t>(atoll(e), 4096);   // tuning knob
        for (int64_t pos = 0; pos < image_bytes;) {
            const int64_t left = image_bytes - pos;
            const int64_t sz = left - uniform < uniform / 2 ? left : uniform;
            pos += sz;
            seg_end.push_back(pos);
        }
    }
    const int64_t n_seg = (int64_t)seg_end.size();
    while ((int64_t)ctx.seg_events.size() < n_seg) {
        cudaEvent_t ev;
        AGX_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx.seg_events.push_back(ev);
    }
    // pinned image: the copies are truly asynchronous, queue them all; pageable image: cudaMemcpyAsync
    // stages through the driver and blocks the host, so stay one segment ahead of the region being scored
    cudaPointerAttributes img_attr;
    const bool img_pinned = cudaPointerGetAttributes(&img_attr, image) == cudaSuccess && img_attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    cudaPointerAttributes out_attr;
    const bool out_pinned = cudaPointerGetAttributes(&out_attr, scores_out) == cudaSuccess && out_attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    int64_t queued = 0;
    auto upload_through = [&](int64_t k_last) -> int {
        for (; queued <= k_last && queued < n_seg; ++queued) {
            const int64_t b = queued ? seg_end[queued - 1] : 0, e = seg_end[queued];
            AGX_CUDA(cudaMemcpyAsync(L.bytes.as<uint8_t>() + b, image + b, (size_t)(e - b), cudaMemcpyHostToDevice, copy_st));
            AGX_CUDA(cudaEventRecord(ctx.seg_events[queued], copy_st));
        }
        return AGX_OK;
    };
    if ((rc = upload_through(img_pinned ? n_seg - 1 : 0)) != AGX_OK) return rc;

    // ---- chunk + score region by region, alternating between the two lanes ------------------------
    // (a lane = stream + workspaces; while the host waits for region k+1's chunk counts on one lane, the
    // DP kernels of region k keep the GPU busy on the other)
    int64_t begin = hlen, remaining = 2 * want_pairs, pairs_done = 0;
    int64_t n_chunks = 0, last_off = -1;
    int32_t last_len = 0;
    bool tail_odd = false;
    int64_t region = 0;
    const bool trace = getenv("AGX_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto now_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    for (int64_t k = 0; k < n_seg && remaining > 0; ++k) {
        if ((rc = upload_through(k + 1)) != AGX_OK) break;
        const int64_t avail = seg_end[k];
        int64_t end = avail;
        if (k + 1 < n_seg) {
            // the region ends after the last newline that has landed
            const void *r = avail > begin ? memrchr(image + begin, '\n', (size_t)(avail - begin)) : nullptr;
            if (!r) continue;                                    // one line spans this whole segment
            end = (const uint8_t *)r - image + 1;
        }
        if (end <= begin) continue;
        const int li = (int)(region++ & 1);
        cudaStream_t st = ctx.lane[li].st, prep = ctx.prep_stream;
        // chunking and the length-class pass run on the high-priority stream, so they start when the
        // segment lands instead of queueing behind the other lane's DP blocks; the lane's buffers are
        // free once its previous region (two regions back) has been scored
        AGX_CUDA(cudaStreamWaitEvent(prep, ctx.seg_events[k], 0));
        if (region > 2) AGX_CUDA(cudaStreamWaitEvent(prep, ctx.lane_done[li], 0));
        int64_t *d_off = nullptr;
        int32_t *d_len = nullptr;
        const double t_p0 = trace ? now_ms() : 0;
        rc = sw_parse_device(ctx.parse[li], L.bytes.as<uint8_t>(), begin, end, line_buf, remaining, image[end - 1],
                             &d_off, &d_len, &n_chunks, &last_off, &last_len, prep);
        if (rc != AGX_OK) break;
        const double t_p1 = trace ? now_ms() : 0;
        const int64_t m = n_chunks / 2;
        if (pairs_done + m > scores_cap) {
            rc = fail(AGX_ERANGE, "sw: scores_out holds " + std::to_string(scores_cap) + " scores, the file has more pairs");
            break;
        }
        if (m > 0) {
            rc = sw_run_device(ctx.lane[li].ws, L.bytes.as<uint8_t>(), d_off, d_len, m, sc,
                               L.out.as<int32_t>() + pairs_done, st, prep);
            if (rc != AGX_OK) break;
            // pinned result array: every region's scores go home behind its own DP kernels
            if (out_pinned)
                AGX_CUDA(cudaMemcpyAsync(scores_out + pairs_done, L.out.as<int32_t>() + pairs_done, (size_t)m * sizeof(int32_t),
                                         cudaMemcpyDeviceToHost, st));
        }
        AGX_CUDA(cudaEventRecord(ctx.lane_done[li], st));
        if (trace)
            fprintf(stderr, "[agx] region %lld seg %lld bytes [%lld, %lld) lane %d: parse %.3f -> %.3f ms, scored %lld pairs by %.3f ms\n",
                    (long long)region - 1, (long long)k, (long long)begin, (long long)end, li, t_p0, t_p1, (long long)m, now_ms());
        pairs_done += m;
        remaining -= 2 * m;
        tail_odd = (n_chunks & 1) != 0;
        // an unpaired last chunk is read again as the first chunk of the next region
        begin = tail_odd ? last_off : end;
    }
    cudaStream_t st = ctx.lane[0].st;
    if (rc != AGX_OK) {
        cudaStreamSynchronize(copy_st);
        cudaStreamSynchronize(ctx.prep_stream);
        cudaStreamSynchronize(ctx.lane[1].st);
        cudaStreamSynchronize(st);
        return rc;
    }
    AGX_CUDA(cudaStreamSynchronize(ctx.lane[1].st));
    if (pairs_done > 0) {
        if (out_pinned) {
            AGX_CUDA(cudaStreamSynchronize(st));           // (copied region by region)
        } else {
            if ((rc = L.h_out.reserve((size_t)pairs_done * sizeof(int32_t))) != AGX_OK) return rc;
            AGX_CUDA(cudaMemcpyAsync(L.h_out.p, L.out.p, (size_t)pairs_done * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            AGX_CUDA(cudaStreamSynchronize(st));
            memcpy(scores_out, L.h_out.p, (size_t)pairs_done * sizeof(int32_t));
        }
    }
    AGX_CUDA(cudaStreamSynchronize(copy_st));
    AGX_CUDA(cudaStreamSynchronize(st));
    if (trace) fprintf(stderr, "[agx] file image done at %.3f ms\n", now_ms());
    if (remaining > 0 && tail_odd) {
        // EOF in the middle of a pair: the reference echoes the dangling first line (:223-227)
        if (dangling_off) *dangling_off = last_off;
        if (dangling_len) *dangling_len = last_len;
    }
    if (n_pairs_out) *n_pairs_out = pairs_done;
    return AGX_OK;
}

int sw_score_batch(const uint8_t *const *a, const int32_t *a_len, const uint8_t *const *b,
                   const int32_t *b_len, int64_t n_pairs, int32_t match, int32_t mismatch,
                   int32_t gap_open, int32_t gap_extend, int32_t *scores_out)
{
    if (n_pairs < 0) return fail(AGX_EINVAL, "sw: n_pairs < 0");
    if (n_pairs == 0) return AGX_OK;
    if (!a || !a_len || !b || !b_len || !scores_out) return fail(AGX_EINVAL, "sw: null argument");
    // gather the pointer arrays into one flat image (a0 b0 a1 b1 ...) in pinned memory, several host threads
    int rc = require_init();
    if (rc != AGX_OK) return rc;
    std::vector<int64_t> off((size_t)n_pairs * 2);
    std::vector<int32_t> len((size_t)n_pairs * 2);
    int64_t total = 0;
    for (int64_t p = 0; p < n_pairs; ++p) {
        if (a_len[p] < 0 || b_len[p] < 0 || (a_len[p] && !a[p]) || (b_len[p] && !b[p]))
            return fail(AGX_EINVAL, "sw: bad sequence " + std::to_string(p));
        off[2 * p] = total; len[2 * p] = a_len[p];
        total += a_len[p];
        off[2 * p + 1] = total; len[2 * p + 1] = b_len[p];
        total += b_len[p];
    }
    DeviceCtx &c0 = *g_ctx[0];
    AGX_CUDA(cudaSetDevice(c0.device));
    if ((rc = c0.h_a.reserve((size_t)total + 1)) != AGX_OK) return rc;
    uint8_t *flat = c0.h_a.as<uint8_t>();
    auto gather = [&](int64_t p0, int64_t p1) {
        for (int64_t p = p0; p < p1; ++p) {
            if (a_len[p]) memcpy(flat + off[2 * p], a[p], (size_t)a_len[p]);
            if (b_len[p]) memcpy(flat + off[2 * p + 1], b[p], (size_t)b_len[p]);
        }
    };
    int n_thr = (int)std::min<int64_t>(std::max(1u, std::min(16u, std::thread::hardware_concurrency())), total / (4 << 20) + 1);
    if (n_thr <= 1) {
        gather(0, n_pairs);
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < n_thr; ++k) th.emplace_back(gather, n_pairs * k / n_thr, n_pairs * (k + 1) / n_thr);
        for (auto &t : th) t.join();
    }
    return sw_flat_impl(flat, total, off.data(), len.data(), n_pairs,
                        SwScoring{match, mismatch, gap_open, gap_extend}, scores_out);
}

int sw_score_batch_device(int32_t device, const uint8_t *d_seqs, int64_t seqs_bytes, const int64_t *d_off,
                          const int32_t *d_len, int64_t n_pairs, int32_t match, int32_t mismatch,
                          int32_t gap_open, int32_t gap_extend, int32_t *d_scores_out, void *stream)
{
    (void)seqs_bytes;
    if (n_pairs < 0) return fail(AGX_EINVAL, "sw: n_pairs < 0");
    if (n_pairs == 0) return AGX_OK;
    if (!d_seqs || !d_off || !d_len || !d_scores_out) return fail(AGX_EINVAL, "sw: null argument");
    DeviceCtx *c = ctx_for_device(device);
    if (!c) return fail(AGX_ENODEVICE, "sw_score_batch_device: device " + std::to_string(device) +
                                           " was not passed to agx_init");
    AGX_CUDA(cudaSetDevice(device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    return sw_run_device(c->sw, d_seqs, d_off, d_len, n_pairs, SwScoring{match, mismatch, gap_open, gap_extend},
                         d_scores_out, st);
}

// Device-resident alignment entry points (one GPU, everything a device pointer on `device`)
int sw_ends_batch_device(int32_t device, const uint8_t *d_seqs, int64_t seqs_bytes, const int64_t *d_off,
                         const int32_t *d_len, int64_t n_pairs, int32_t match, int32_t mismatch, int32_t gap_open,
                         int32_t gap_extend, int32_t *d_scores_out, int32_t *d_ends_out, void *stream)
{
    (void)seqs_bytes;
    if (n_pairs < 0) return fail(AGX_EINVAL, "sw ends: n_pairs < 0");
    if (n_pairs == 0) return AGX_OK;
    if (!d_seqs || !d_off || !d_len || !d_scores_out || !d_ends_out) return fail(AGX_EINVAL, "sw ends: null argument");
    DeviceCtx *c = ctx_for_device(device);
    if (!c) return fail(AGX_ENODEVICE, "sw_ends_batch_device: device " + std::to_string(device) + " was not passed to agx_init");
    AGX_CUDA(cudaSetDevice(device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    return sw_align_run_device(c->align, d_seqs, d_off, d_len, n_pairs, SwScoring{match, mismatch, gap_open, gap_extend}, 1,
                               0, d_scores_out, d_ends_out, nullptr, nullptr, st);
}

int sw_align_batch_device(int32_t device, const uint8_t *d_seqs, int64_t seqs_bytes, const int64_t *d_off,
                          const int32_t *d_len, int64_t n_pairs, int32_t match, int32_t mismatch, int32_t gap_open,
                          int32_t gap_extend, int32_t *d_scores_out, int32_t *d_coords_out, int64_t *d_cigar_off_out,
                          uint32_t *d_cigar_out, int64_t cigar_cap, int64_t *cigar_total_out, void *stream)
{
    (void)seqs_bytes;
    if (cigar_total_out) *cigar_total_out = 0;
    if (n_pairs < 0) return fail(AGX_EINVAL, "sw align: n_pairs < 0");
    if (n_pairs == 0) return AGX_OK;
    if (!d_seqs || !d_off || !d_len || !d_scores_out || !d_coords_out || !d_cigar_off_out || (!d_cigar_out && cigar_cap > 0) ||
        cigar_cap < 0)
        return fail(AGX_EINVAL, "sw align: null argument");
    DeviceCtx *c = ctx_for_device(device);
    if (!c) return fail(AGX_ENODEVICE, "sw_align_batch_device: device " + std::to_string(device) + " was not passed to agx_init");
    AGX_CUDA(cudaSetDevice(device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    // one chunk: the score matrices of the whole batch must fit what is free (else AGX_ENOMEM: cut the batch)
    int rc;
    if ((rc = c->al[0].ends.reserve((size_t)n_pairs * 2 * sizeof(int32_t))) != AGX_OK) return rc;
    int64_t total = 0, avail = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if ((rc = align_memory(*c, attempt == 1, &avail)) != AGX_OK) return rc;
        rc = sw_align_run_device(c->align, d_seqs, d_off, d_len, n_pairs, SwScoring{match, mismatch, gap_open, gap_extend}, 2,
                                 avail - ((int64_t)1 << 30), d_scores_out, c->al[0].ends.as<int32_t>(), d_coords_out, &total, st);
        if (rc != AGX_ENOMEM) break;            // out of memory against a remembered figure: ask again once
    }
    if (rc != AGX_OK) return rc;
    AGX_CUDA(cudaMemcpyAsync(d_cigar_off_out, c->align.cig_off, (size_t)(n_pairs + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    if (cigar_total_out) *cigar_total_out = total;
    if (total > cigar_cap)
        return fail(AGX_ERANGE, "sw align: " + std::to_string(total) + " CIGAR runs do not fit cigar_cap = " + std::to_string(cigar_cap) +
                                    " (scores, coordinates and offsets are complete)");
    return sw_align_gather_device(c->align, n_pairs, d_cigar_out, st);
}

int sw_score_shards_device(const agx_sw_shard *shards, int32_t n_shards, int32_t match, int32_t mismatch,
                           int32_t gap_open, int32_t gap_extend)
{
    const SwScoring sc{match, mismatch, gap_open, gap_extend};
    return for_each_shard(shards, n_shards, [&](DeviceCtx &c, const agx_sw_shard &s) -> int {
        if (s.n_pairs < 0) return fail(AGX_EINVAL, "sw: n_pairs < 0");
        if (s.n_pairs == 0) return AGX_OK;
        if (!s.d_seqs || !s.d_off || !s.d_len || !s.d_scores_out) return fail(AGX_EINVAL, "sw: null argument");
        return sw_run_device(c.sw, s.d_seqs, s.d_off, s.d_len, s.n_pairs, sc, s.d_scores_out, c.stream);
    });
}

int pairhmm_forward_shards_device(const agx_hmm_shard *shards, int32_t n_shards, int32_t fp64_rescue)
{
    return for_each_shard(shards, n_shards, [&](DeviceCtx &c, const agx_hmm_shard &s) -> int {
        if (s.n_reads < 0 || s.n_haps < 0 || s.n_batches < 0 || s.n_pairs < 0) return fail(AGX_EINVAL, "pairhmm: negative count");
        if (s.n_reads == 0 || s.n_pairs == 0) return AGX_OK;
        if (!s.d_buf || !s.d_read_field_off || !s.d_read_len || !s.d_read_batch || !s.d_read_out_off || !s.d_hap_off ||
            !s.d_hap_len || !s.d_batch_hap_start || !s.d_log10_out)
            return fail(AGX_EINVAL, "pairhmm: null argument");
        HmmBatchView v;
        v.buf = s.d_buf; v.read_field_off = s.d_read_field_off; v.read_len = s.d_read_len; v.read_batch = s.d_read_batch;
        v.n_reads = s.n_reads; v.hap_off = s.d_hap_off; v.hap_len = s.d_hap_len; v.n_haps = s.n_haps;
        v.batch_hap_start = s.d_batch_hap_start; v.n_batches = s.n_batches;
        return hmm_run_device(c.hmm, v, s.buf_bytes, s.d_read_out_off, s.n_pairs, g_gatk.load(), g_force64.load() != 0,
                              fp64_rescue != 0, s.d_log10_out, c.stream);
    });
}

int64_t agx_pairhmm_rescue_count(int32_t device)
{
    DeviceCtx *c = ctx_for_device(device);
    return c ? c->hmm.last_rescue : -1;
}

int pairhmm_forward_batches_flat(const uint8_t *buf, int64_t buf_bytes, const int64_t *read_field_off,
                                 const int32_t *read_len, int64_t n_reads, const int64_t *hap_off,
                                 const int32_t *hap_len, int64_t n_haps, const int64_t *batch_read_start,
                                 const int64_t *batch_hap_start, int64_t n_batches, double *log10_out)
{
    HmmHost h;
    h.buf = buf; h.buf_bytes = buf_bytes;
    h.read_field_off = read_field_off; h.read_len = read_len; h.n_reads = n_reads;
    h.hap_off = hap_off; h.hap_len = hap_len; h.n_haps = n_haps;
    h.batch_read_start = batch_read_start; h.batch_hap_start = batch_hap_start; h.n_batches = n_batches;
    h.n_pairs = 0;
    return hmm_flat_impl(h, log10_out);
}

// What one GPU did with (its range of) a PairHMM file image: results stay in c.d_out on the device.
struct HmmImageRun {
    int64_t n_out = 0;              // results in c.d_out
    std::vector<int32_t> bp;        // reads x haplotypes of every complete batch
    int32_t incomplete = 0;         // the range ends inside a batch (see agx.h)
    int64_t parsed_end = 0;         // where parsing stopped: the range's size when every byte belongs to a complete batch
};

// Like sw_score_file_image: the image is uploaded in segments on the copy stream; each region (whole batches)
// is parsed and paired on the high-priority stream and scored on one of two alternating lanes, so the upload
// and the host round trips of region k+1 hide behind the stream kernels of region k.  A batch that straddles
// a region boundary is parsed again at the start of the next region.  All streams are idle on return.
static int hmm_image_run(DeviceCtx &c, const uint8_t *image, int64_t image_bytes, HmmImageRun &run)
{
    int rc;
    cudaStream_t lane_st[2] = {c.lane[0].st, c.lane[1].st};
    HmmWorkspace *lane_ws[2] = {&c.hmm, &c.hmm_b};
    HmmParseWorkspace *lane_parse[2] = {&c.hmm_parse, &c.hmm_parse_b};
    if ((rc = c.d_bytes.reserve((size_t)image_bytes + 64)) != AGX_OK) return rc;

    // segments: a short first one so that scoring starts early, then doubling (the stream kernels need ~7x
    // the time the upload of the same bytes takes, so the copy never falls behind)
    std::vector<int64_t> seg_end;
    {
        int64_t fixed = 0;
        if (const char *e = getenv("AGX_HMM_IMAGE_SEGMENT")) fixed = atoll(e);   // tuning knob: fixed segment bytes
        int64_t sz = (int64_t)16 << 20;
        for (int64_t pos = 0; pos < image_bytes;) {
            int64_t step = fixed > 0 ? fixed : sz;
            if (image_bytes - pos - step < step / 2) step = image_bytes - pos;
            pos += step;
            seg_end.push_back(pos);
            sz *= 2;
        }
    }
    const int64_t n_seg = (int64_t)seg_end.size();
    while ((int64_t)c.seg_events.size() < n_seg) {
        cudaEvent_t ev;
        AGX_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        c.seg_events.push_back(ev);
    }
    cudaPointerAttributes img_attr;
    const bool img_pinned = cudaPointerGetAttributes(&img_attr, image) == cudaSuccess && img_attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    int64_t queued = 0;
    auto upload_through = [&](int64_t k_last) -> int {
        for (; queued <= k_last && queued < n_seg; ++queued) {
            const int64_t b = queued ? seg_end[queued - 1] : 0, e = seg_end[queued];
            AGX_CUDA(cudaMemcpyAsync(c.d_bytes.as<uint8_t>() + b, image + b, (size_t)(e - b), cudaMemcpyHostToDevice, c.copy_stream));
            AGX_CUDA(cudaEventRecord(c.seg_events[queued], c.copy_stream));
        }
        return AGX_OK;
    };
    if ((rc = upload_through(img_pinned ? n_seg - 1 : 0)) != AGX_OK) return rc;

    // results: the number of outputs is not known before the last region is parsed, so the device buffer
    // grows region by region (regions are few); batch counts go to the host as they are parsed
    std::vector<int32_t> &bp_all = run.bp;
    bp_all.clear();
    int64_t out_done = 0, begin = 0, region = 0;
    int32_t inc = 0;
    int32_t carry_nr = 0, carry_nh = 0;       // header counts of the batch before `begin` (antidiagsPairHMM.c:345-346)
    cudaStream_t prep = c.prep_stream;
    for (int64_t k = 0; k < n_seg; ++k) {
        if ((rc = upload_through(k + 1)) != AGX_OK) break;
        const int64_t avail = seg_end[k];
        int64_t end = avail;
        if (k + 1 < n_seg) {
            const void *r = avail > begin ? memrchr(image + begin, '\n', (size_t)(avail - begin)) : nullptr;
            if (!r) continue;
            end = (const uint8_t *)r - image + 1;
        }
        if (end <= begin) continue;
        const int li = (int)(region & 1);
        AGX_CUDA(cudaStreamWaitEvent(prep, c.seg_events[k], 0));
        if (region >= 2) AGX_CUDA(cudaStreamWaitEvent(prep, c.lane_done[li], 0));   // the lane's arrays are free again
        HmmParsed ps;
        if ((rc = hmm_parse_device(*lane_parse[li], c.d_bytes.as<uint8_t>(), begin, end, image[end - 1], &ps, prep, carry_nr,
                                   carry_nh)) != AGX_OK)
            break;
        carry_nr = ps.last_nr;
        carry_nh = ps.last_nh;
        const bool last_region = (k + 1 == n_seg);
        if (last_region) inc = ps.incomplete;
        if (ps.n_batches == 0) {
            if (ps.next_begin <= begin && !last_region) continue;      // one batch spans the whole region: wait for more
            begin = ps.next_begin;
            continue;
        }
        ++region;
        // batch counts of this region (pageable destination: the copy is complete when the call returns)
        const size_t b0 = bp_all.size();
        bp_all.resize(b0 + (size_t)ps.n_batches);
        AGX_CUDA(cudaMemcpyAsync(bp_all.data() + b0, ps.batch_pairs, (size_t)ps.n_batches * sizeof(int32_t), cudaMemcpyDeviceToHost, prep));
        AGX_CUDA(cudaStreamSynchronize(prep));
        if (ps.n_out > 0) {
            if ((size_t)(out_done + ps.n_out) * sizeof(double) > c.d_out.cap) {
                // grow, keeping what earlier regions wrote (both lanes must be idle while the buffer moves)
                AGX_CUDA(cudaStreamSynchronize(lane_st[0]));
                AGX_CUDA(cudaStreamSynchronize(lane_st[1]));
                DevBuf bigger;
                const double frac = (double)(end - 0) / (double)image_bytes;
                const size_t want = (size_t)((double)(out_done + ps.n_out) / (frac > 0.05 ? frac : 0.05) * 1.1) * sizeof(double);
                if ((rc = bigger.reserve(std::max(want, (size_t)(out_done + ps.n_out) * sizeof(double)))) != AGX_OK) break;
                if (out_done > 0) AGX_CUDA(cudaMemcpy(bigger.p, c.d_out.p, (size_t)out_done * sizeof(double), cudaMemcpyDeviceToDevice));
                c.d_out.release();
                c.d_out = bigger;
            }
            HmmBatchView v;
            v.buf = c.d_bytes.as<uint8_t>();
            v.read_field_off = ps.read_field_off;
            v.read_len = ps.read_len;
            v.read_batch = ps.read_batch;
            v.n_reads = ps.n_reads;
            v.hap_off = ps.hap_off;
            v.hap_len = ps.hap_len;
            v.n_haps = ps.n_haps;
            v.batch_hap_start = ps.batch_hap_start;
            v.n_batches = ps.n_batches;
            rc = hmm_run_device(*lane_ws[li], v, image_bytes, ps.read_out_off, ps.n_out, g_gatk.load(),
                                g_force64.load() != 0, 2, c.d_out.as<double>() + out_done, lane_st[li], prep);
            if (rc != AGX_OK) break;
            out_done += ps.n_out;
        }
        AGX_CUDA(cudaEventRecord(c.lane_done[li], lane_st[li]));
        begin = ps.next_begin;
    }
    cudaStreamSynchronize(c.copy_stream);
    cudaStreamSynchronize(prep);
    cudaStreamSynchronize(lane_st[0]);
    cudaStreamSynchronize(lane_st[1]);
    if (rc != AGX_OK) return rc;
    AGX_CUDA(cudaGetLastError());
    run.n_out = out_done;
    run.incomplete = inc;
    run.parsed_end = begin;
    return AGX_OK;
}

// Several GPUs: the image is cut at header-shaped lines near the even split points and every GPU runs the
// single-GPU pipeline on its own byte range (one host thread per GPU, no collective).  A cut is a true batch
// boundary exactly when the range before it parses to its last byte without a dangling batch -- by induction
// from the start of the file that is where the reference's own walk reads its next header -- so the ranges are
// checked for that and anything else (irregular headers, a "header" inside a batch) returns 1: the caller then
// takes the single-GPU path, which follows the reference's walk.
static int hmm_file_image_multi(const uint8_t *image, int64_t image_bytes, std::vector<HmmImageRun> &runs, std::vector<int64_t> &cut)
{
    const int n_dev = (int)g_ctx.size();
    if (n_dev < 2 || image_bytes < (int64_t)n_dev * (8 << 20) || getenv("AGX_HMM_IMAGE_ONE_GPU")) return 1;
    auto header_shaped = [&](int64_t s, int64_t e) {         // [s, e): a line without its newline
        int64_t p = s;
        int d0 = 0, bl = 0, d1 = 0;
        while (p < e && image[p] >= '0' && image[p] <= '9') { ++p; ++d0; }
        while (p < e && (image[p] == ' ' || image[p] == '\t')) { ++p; ++bl; }
        while (p < e && image[p] >= '0' && image[p] <= '9') { ++p; ++d1; }
        return p == e && d0 >= 1 && d0 <= 9 && bl >= 1 && d1 >= 1 && d1 <= 9;
    };
    cut.assign(n_dev + 1, image_bytes);
    cut[0] = 0;
    for (int d = 1; d < n_dev; ++d) {
        const int64_t nominal = image_bytes * d / n_dev;
        const void *nl = memchr(image + nominal, '\n', (size_t)(image_bytes - nominal));
        int64_t s = nl ? (const uint8_t *)nl - image + 1 : image_bytes;
        int64_t found = -1;
        for (int lines = 0; s < image_bytes && lines < (1 << 16); ++lines) {
            const void *e = memchr(image + s, '\n', (size_t)(image_bytes - s));
            const int64_t le = e ? (const uint8_t *)e - image : image_bytes;
            if (header_shaped(s, le)) { found = s; break; }
            s = le + 1;
        }
        if (found < 0 || found <= cut[d - 1]) return 1;
        cut[d] = found;
    }
    runs.assign(n_dev, HmmImageRun());
    int rc = for_each_device(n_dev, [&](DeviceCtx &c, int d) -> int {
        return hmm_image_run(c, image + cut[d], cut[d + 1] - cut[d], runs[d]);
    });
    if (rc != AGX_OK) return rc;
    for (int d = 0; d + 1 < n_dev; ++d)
        if (runs[d].incomplete != 0 || runs[d].parsed_end != cut[d + 1] - cut[d]) return 1;
    return AGX_OK;
}

int pairhmm_forward_file_image(const uint8_t *image, int64_t image_bytes, const double **log10_out, int64_t *n_out,
                               const int32_t **batch_pairs, int64_t *n_batches, int32_t *incomplete)
{
    if (log10_out) *log10_out = nullptr;
    if (n_out) *n_out = 0;
    if (batch_pairs) *batch_pairs = nullptr;
    if (n_batches) *n_batches = 0;
    if (incomplete) *incomplete = 0;
    if (image_bytes < 0 || (image_bytes > 0 && !image)) return fail(AGX_EINVAL, "pairhmm: null file image");
    if (!log10_out || !n_out || !batch_pairs || !n_batches) return fail(AGX_EINVAL, "pairhmm: null argument");
    if (image_bytes == 0) return AGX_OK;
    int rc = require_init();
    if (rc != AGX_OK) return rc;
    DeviceCtx &c0 = *g_ctx[0];

    std::vector<HmmImageRun> runs;
    std::vector<int64_t> cut;
    rc = hmm_file_image_multi(image, image_bytes, runs, cut);
    if (rc < 0) return rc;
    if (rc == 1) {
        // one GPU (or an image the ranges could not be validated on)
        runs.assign(1, HmmImageRun());
        AGX_CUDA(cudaSetDevice(c0.device));
        if ((rc = hmm_image_run(c0, image, image_bytes, runs[0])) != AGX_OK) return rc;
    }
    // gather: every GPU copies its results to their place in one pinned array; batch counts are concatenated
    const int n_used = (int)runs.size();
    std::vector<int64_t> out_at(n_used + 1, 0);
    size_t n_bp = 0;
    for (int d = 0; d < n_used; ++d) { out_at[d + 1] = out_at[d] + runs[d].n_out; n_bp += runs[d].bp.size(); }
    const int64_t total = out_at[n_used];
    AGX_CUDA(cudaSetDevice(c0.device));
    if (total > 0 && (rc = c0.h_out.reserve((size_t)total * sizeof(double))) != AGX_OK) return rc;
    if (n_bp > 0 && (rc = c0.h_b.reserve(n_bp * sizeof(int32_t))) != AGX_OK) return rc;
    double *h_all = c0.h_out.as<double>();
    rc = for_each_device(n_used, [&](DeviceCtx &c, int d) -> int {
        if (runs[d].n_out == 0) return AGX_OK;
        AGX_CUDA(cudaMemcpyAsync(h_all + out_at[d], c.d_out.p, (size_t)runs[d].n_out * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        AGX_CUDA(cudaStreamSynchronize(c.stream));
        return AGX_OK;
    });
    if (rc != AGX_OK) return rc;
    size_t w = 0;
    for (int d = 0; d < n_used; ++d) {
        if (!runs[d].bp.empty()) memcpy(c0.h_b.as<int32_t>() + w, runs[d].bp.data(), runs[d].bp.size() * sizeof(int32_t));
        w += runs[d].bp.size();
    }
    if (incomplete) *incomplete = runs[n_used - 1].incomplete;
    *n_out = total;
    *n_batches = (int64_t)n_bp;
    *log10_out = total > 0 ? h_all : nullptr;
    *batch_pairs = n_bp ? c0.h_b.as<int32_t>() : nullptr;
    return AGX_OK;
}

int pairhmm_forward_batch(int32_t n_reads, const uint8_t *const *bases, const uint8_t *const *q,
                          const uint8_t *const *qi, const uint8_t *const *qd, const uint8_t *const *qg,
                          const int32_t *read_len, int32_t n_haps, const uint8_t *const *haps,
                          const int32_t *hap_len, double *log10_out)
{
    if (n_reads < 0 || n_haps < 0) return fail(AGX_EINVAL, "pairhmm: negative count");
    if (n_reads == 0 || n_haps == 0) return AGX_OK;
    if (!bases || !q || !qi || !qd || !qg || !read_len || !haps || !hap_len || !log10_out)
        return fail(AGX_EINVAL, "pairhmm: null argument");
    int64_t total = 0;
    for (int r = 0; r < n_reads; ++r) {
        if (read_len[r] < 1) return fail(AGX_ERANGE, "pairhmm: read " + std::to_string(r) + " is empty");
        if (!bases[r] || !q[r] || !qi[r] || !qd[r] || !qg[r]) return fail(AGX_EINVAL, "pairhmm: null read field");
        total += 5 * (int64_t)read_len[r];
    }
    for (int x = 0; x < n_haps; ++x) {
        if (hap_len[x] < 1) return fail(AGX_ERANGE, "pairhmm: haplotype " + std::to_string(x) + " is empty");
        if (!haps[x]) return fail(AGX_EINVAL, "pairhmm: null haplotype");
        total += hap_len[x];
    }
    std::vector<uint8_t> flat((size_t)total + 1);
    std::vector<int64_t> rfo((size_t)n_reads * 5), ho((size_t)n_haps);
    int64_t w = 0;
    for (int r = 0; r < n_reads; ++r) {
        const uint8_t *f[5] = {bases[r], q[r], qi[r], qd[r], qg[r]};
        for (int k = 0; k < 5; ++k) {
            rfo[5 * (size_t)r + k] = w;
            memcpy(flat.data() + w, f[k], (size_t)read_len[r]);
            w += read_len[r];
        }
    }
    for (int x = 0; x < n_haps; ++x) {
        ho[x] = w;
        memcpy(flat.data() + w, haps[x], (size_t)hap_len[x]);
        w += hap_len[x];
    }
    const int64_t brs[2] = {0, n_reads}, bhs[2] = {0, n_haps};
    return pairhmm_forward_batches_flat(flat.data(), total, rfo.data(), read_len, n_reads, ho.data(), hap_len,
                                        n_haps, brs, bhs, 1, log10_out);
}

int pairhmm_forward_batches_device(int32_t device, const uint8_t *d_buf, int64_t buf_bytes,
                                   const int64_t *d_read_field_off, const int32_t *d_read_len,
                                   const int32_t *d_read_batch, const int64_t *d_read_out_off,
                                   int64_t n_reads, const int64_t *d_hap_off, const int32_t *d_hap_len,
                                   int64_t n_haps, const int64_t *d_batch_hap_start, int64_t n_batches,
                                   int64_t n_pairs, int32_t fp64_rescue, double *d_log10_out, void *stream)
{
    if (n_reads < 0 || n_haps < 0 || n_batches < 0 || n_pairs < 0 || buf_bytes < 0) return fail(AGX_EINVAL, "pairhmm: negative count");
    if (n_reads == 0 || n_pairs == 0) return AGX_OK;
    if (!d_buf || !d_read_field_off || !d_read_len || !d_read_batch || !d_read_out_off || !d_hap_off ||
        !d_hap_len || !d_batch_hap_start || !d_log10_out)
        return fail(AGX_EINVAL, "pairhmm: null argument");
    DeviceCtx *c = ctx_for_device(device);
    if (!c) return fail(AGX_ENODEVICE, "pairhmm_forward_batches_device: device " + std::to_string(device) +
                                           " was not passed to agx_init");
    AGX_CUDA(cudaSetDevice(device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    HmmBatchView v;
    v.buf = d_buf;
    v.read_field_off = d_read_field_off;
    v.read_len = d_read_len;
    v.read_batch = d_read_batch;
    v.n_reads = n_reads;
    v.hap_off = d_hap_off;
    v.hap_len = d_hap_len;
    v.n_haps = n_haps;
    v.batch_hap_start = d_batch_hap_start;
    v.n_batches = n_batches;
    return hmm_run_device(c->hmm, v, buf_bytes, d_read_out_off, n_pairs, g_gatk.load(), g_force64.load() != 0,
                          fp64_rescue != 0, d_log10_out, st);
}

}  // extern "C"
