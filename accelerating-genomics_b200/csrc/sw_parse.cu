// sw_parse.cu -- device-side line splitter for generator.py-style Smith-Waterman files.
//
// The reference reads its input with fgets() into a MAX_LINE_LENGTH (1000) byte buffer
// (smithWaterman/antidiagonalSmithWaterman.c:201-227): every sequence is one fgets() chunk -- a line
// including its '\n', or, for a line of >= MAX_LINE_LENGTH bytes, successive pieces of
// MAX_LINE_LENGTH-1 bytes.  This file reproduces that chunking ON THE GPU from the raw file image, so a
// driver uploads the image once and never builds (offset, length) arrays on the host:
//
//   nl_count_kernel   128-bit coalesced loads, newlines per 4 KiB tile            (HBM-bound)
//   scan (3 kernels)  exclusive prefix sums, reused for tiles and for chunks-per-line
//   nl_emit_kernel    position of every newline, in order                          (HBM-bound)
//   line_chunks_kernel / chunk_emit_kernel   chunks per line, then (off, len) of every chunk
//
// The DP kernels (sw_kernels.cu) then consume the chunk table directly.
#include <algorithm>

#include "common.cuh"

namespace agx {

namespace {

constexpr int TILE_THREADS = 256;
constexpr int TILE_BYTES = TILE_THREADS * 16;

__device__ __forceinline__ uint32_t nl_mask16(const uint8_t *p, int64_t pos, int64_t begin, int64_t end)
{
    // bit i set <=> byte pos+i is '\n' and begin <= pos+i < end.  pos is a multiple of 16 and the image
    // buffer is 256-byte aligned with slack behind it, so the 128-bit load is always legal.
    uint32_t m = 0;
    if (pos >= begin && pos + 16 <= end) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p + pos);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t x = w[k] ^ 0x0a0a0a0au;                       // zero byte where '\n'
            const uint32_t z = ((x - 0x01010101u) & ~x & 0x80808080u);    // 0x80 in every zero byte
#pragma unroll
            for (int b = 0; b < 4; ++b) m |= ((z >> (8 * b + 7)) & 1u) << (4 * k + b);
        }
    } else {
        for (int i = 0; i < 16; ++i)
            if (pos + i >= begin && pos + i < end && p[pos + i] == '\n') m |= 1u << i;
    }
    return m;
}

__global__ void __launch_bounds__(TILE_THREADS)
nl_count_kernel(const uint8_t *__restrict__ img, int64_t begin, int64_t end, int32_t *__restrict__ tile_count)
{
    __shared__ int32_t s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const int64_t pos = (begin & ~(int64_t)15) + (int64_t)blockIdx.x * TILE_BYTES + threadIdx.x * 16;
    int c = pos < end ? __popc(nl_mask16(img, pos, begin, end)) : 0;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_sum, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_count[blockIdx.x] = s_sum;
}

__global__ void __launch_bounds__(TILE_THREADS)
nl_emit_kernel(const uint8_t *__restrict__ img, int64_t begin, int64_t end,
               const int64_t *__restrict__ tile_base, int64_t *__restrict__ nl_pos)
{
    __shared__ int32_t s_warp[TILE_THREADS / 32];
    const int64_t pos = (begin & ~(int64_t)15) + (int64_t)blockIdx.x * TILE_BYTES + threadIdx.x * 16;
    const uint32_t mask = pos < end ? nl_mask16(img, pos, begin, end) : 0u;
    const int c = __popc(mask);
    // exclusive scan of c over the block
    int incl = c;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    int64_t base = tile_base[blockIdx.x];
    for (int k = 0; k < w; ++k) base += s_warp[k];
    int64_t idx = base + incl - c;
    uint32_t m = mask;
    while (m) {
        const int b = __ffs(m) - 1;
        nl_pos[idx++] = pos + b;
        m &= m - 1;
    }
}

// ---- exclusive scan of int32 (three small kernels; n up to 2^31) -------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ tile_sum)
{
    __shared__ int64_t s[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int64_t v = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) v += in[base + i];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int k = 0; k < SCAN_THREADS / 32; ++k) t += s[k];
        tile_sum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
scan_tiles_kernel(int64_t *__restrict__ tile_sum, int64_t n_tiles, int64_t *__restrict__ total)
{
    // one block walks the tile sums in chunks of 1024 (n_tiles is n / 1024: small)
    __shared__ int64_t s[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t b = 0; b < n_tiles; b += 1024) {
        const int64_t i = b + threadIdx.x;
        const int64_t v = i < n_tiles ? tile_sum[i] : 0;
        int64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) s[w] = incl;
        __syncthreads();
        int64_t off = carry;
        for (int k = 0; k < w; ++k) off += s[k];
        if (i < n_tiles) tile_sum[i] = off + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const int32_t *__restrict__ in, int64_t n, const int64_t *__restrict__ tile_off,
                  int64_t *__restrict__ out)
{
    __shared__ int64_t s[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int32_t x[SCAN_ITEMS];
    int64_t v = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        x[i] = (base + i < n) ? in[base + i] : 0;
        v += x[i];
    }
    int64_t incl = v;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t u = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += u;
    }
    if (lane == 31) s[w] = incl;
    __syncthreads();
    int64_t off = tile_off[blockIdx.x];
    for (int k = 0; k < w; ++k) off += s[k];
    off += incl - v;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = off;
        off += x[i];
    }
}

// exclusive scan: out[i] = sum_{k<i} in[k]; *d_total = sum of all.  tmp holds ceil(n / SCAN_TILE) int64.
int exclusive_scan(const int32_t *in, int64_t n, int64_t *out, int64_t *tmp, int64_t *d_total, cudaStream_t st)
{
    if (n <= 0) {
        AGX_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int64_t), st));
        return AGX_OK;
    }
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_reduce_kernel<<<(int)tiles, SCAN_THREADS, 0, st>>>(in, n, tmp);
    scan_tiles_kernel<<<1, 1024, 0, st>>>(tmp, tiles, d_total);
    scan_apply_kernel<<<(int)tiles, SCAN_THREADS, 0, st>>>(in, n, tmp, out);
    count_launch(3);
    AGX_CUDA(cudaGetLastError());
    return AGX_OK;
}

// line k = bytes (prev newline, this newline]; the tail after the last newline is one more line
__global__ void __launch_bounds__(256)
line_chunks_kernel(const int64_t *__restrict__ nl_pos, int64_t n_nl, int64_t begin, int64_t end, int32_t cap,
                   int64_t n_lines, int32_t *__restrict__ chunks_per_line)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines) return;
    const int64_t s = k == 0 ? begin : nl_pos[k - 1] + 1;
    const int64_t e = k < n_nl ? nl_pos[k] + 1 : end;
    const int64_t L = e - s;
    chunks_per_line[k] = (int32_t)((L + cap - 1) / cap);
}

__global__ void __launch_bounds__(256)
chunk_emit_kernel(const int64_t *__restrict__ nl_pos, int64_t n_nl, int64_t begin, int64_t end, int32_t cap,
                  int64_t n_lines, const int64_t *__restrict__ chunk_base, int64_t max_chunks,
                  int64_t *__restrict__ off, int32_t *__restrict__ len, int64_t *__restrict__ totals)
{
    // totals[0] = number of chunks in the region (written by the scan); totals[1], totals[2] receive the
    // offset and length of the last chunk that is kept
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines) return;
    int64_t s = k == 0 ? begin : nl_pos[k - 1] + 1;
    const int64_t e = k < n_nl ? nl_pos[k] + 1 : end;
    int64_t c = chunk_base[k];
    const int64_t last = (totals[0] < max_chunks ? totals[0] : max_chunks) - 1;
    while (s < e && c < max_chunks) {
        const int64_t l = (e - s) < cap ? (e - s) : cap;
        off[c] = s;
        len[c] = (int32_t)l;
        if (c == last) { totals[1] = s; totals[2] = l; }
        s += l;
        ++c;
    }
}

}  // namespace

int device_exclusive_scan(const int32_t *in, int64_t n, int64_t *out, int64_t *tmp, int64_t *d_total, cudaStream_t st)
{
    return exclusive_scan(in, n, out, tmp, d_total, st);
}
int64_t device_scan_tmp_elems(int64_t n) { return n / SCAN_TILE + 2; }

// Position of every '\n' in d_img[begin, end), in order.  *d_nl_pos points into ws.buf2 (n_nl int64 followed
// by `extra_bytes` the caller may use).  Synchronises `st` once (the newline count sizes the table).
int text_newline_index(SwParseWorkspace &ws, const uint8_t *d_img, int64_t begin, int64_t end, int64_t extra_bytes_per_line,
                       int64_t **d_nl_pos, int64_t *n_nl_out, void **d_extra, cudaStream_t st)
{
    *d_nl_pos = nullptr;
    *n_nl_out = 0;
    if (d_extra) *d_extra = nullptr;
    if (end <= begin) return AGX_OK;
    const int64_t tiles = (end - (begin & ~(int64_t)15) + TILE_BYTES - 1) / TILE_BYTES;
    if (!ws.h_total) AGX_CUDA(cudaMallocHost(&ws.h_total, 4 * sizeof(int64_t)));
    auto align = [](int64_t x) { return (x + 255) / 256 * 256; };
    const int64_t sz_tc = align(tiles * 4), sz_tb = align(tiles * 8), sz_tmp = align((tiles / SCAN_TILE + 2) * 8);
    const int64_t need = sz_tc + sz_tb + sz_tmp + 256;
    if (need > ws.cap) {
        if (ws.buf) cudaFree(ws.buf);
        ws.buf = nullptr; ws.cap = 0;
        AGX_CUDA(cudaMalloc(&ws.buf, (size_t)need));
        ws.cap = need;
    }
    uint8_t *base = reinterpret_cast<uint8_t *>(ws.buf);
    int32_t *tile_count = reinterpret_cast<int32_t *>(base);
    int64_t *tile_base64 = reinterpret_cast<int64_t *>(base + sz_tc);
    int64_t *tmp = reinterpret_cast<int64_t *>(base + sz_tc + sz_tb);
    int64_t *d_total = reinterpret_cast<int64_t *>(base + sz_tc + sz_tb + sz_tmp);
    nl_count_kernel<<<(int)tiles, TILE_THREADS, 0, st>>>(d_img, begin, end, tile_count);
    count_launch();
    AGX_CUDA(cudaGetLastError());
    int rc = exclusive_scan(tile_count, tiles, tile_base64, tmp, d_total, st);
    if (rc != AGX_OK) return rc;
    AGX_CUDA(cudaMemcpyAsync(ws.h_total, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    AGX_CUDA(cudaStreamSynchronize(st));
    const int64_t n_nl = ws.h_total[0];
    if (n_nl > (int64_t)1 << 31) return fail(AGX_ERANGE, "more than 2^31 lines");
    const int64_t sz_np = align(std::max<int64_t>(n_nl, 1) * 8);
    const int64_t need2 = sz_np + align((n_nl + 2) * extra_bytes_per_line) + 256;
    if (need2 > ws.cap2) {
        if (ws.buf2) cudaFree(ws.buf2);
        ws.buf2 = nullptr; ws.cap2 = 0;
        AGX_CUDA(cudaMalloc(&ws.buf2, (size_t)need2));
        ws.cap2 = need2;
    }
    int64_t *nl_pos = reinterpret_cast<int64_t *>(ws.buf2);
    nl_emit_kernel<<<(int)tiles, TILE_THREADS, 0, st>>>(d_img, begin, end, tile_base64, nl_pos);
    count_launch();
    AGX_CUDA(cudaGetLastError());
    *d_nl_pos = nl_pos;
    *n_nl_out = n_nl;
    if (d_extra) *d_extra = reinterpret_cast<uint8_t *>(ws.buf2) + sz_np;
    return AGX_OK;
}

void sw_parse_workspace_free(SwParseWorkspace &ws)
{
    if (ws.buf) cudaFree(ws.buf);
    if (ws.buf2) cudaFree(ws.buf2);
    if (ws.h_total) cudaFreeHost(ws.h_total);
    ws = SwParseWorkspace();
}

// Splits d_img[begin, end) into fgets(line_buf) chunks; `begin` must be a position fgets() would start a
// read at (a line start or a chunk start -- fgets keeps no state besides the file position).  On return
// *d_off / *d_len point into the workspace and hold min(total chunks, max_chunks) entries; *n_chunks_out
// is that count and *last_off / *last_len describe the last of them.  last_byte = d_img[end-1] if the
// caller knows it, -1 to have it read back.  Synchronises `st` twice (newline count, chunk count).
int sw_parse_device(SwParseWorkspace &ws, const uint8_t *d_img, int64_t begin, int64_t end, int32_t line_buf,
                    int64_t max_chunks, int last_byte, int64_t **d_off, int32_t **d_len, int64_t *n_chunks_out,
                    int64_t *last_off, int32_t *last_len, cudaStream_t st)
{
    *n_chunks_out = 0;
    *d_off = nullptr;
    *d_len = nullptr;
    if (last_off) *last_off = -1;
    if (last_len) *last_len = 0;
    if (end <= begin || max_chunks <= 0) return AGX_OK;
    if (line_buf < 2) return fail(AGX_EINVAL, "sw: line buffer must hold at least one character");
    const int32_t cap = line_buf - 1;
    const int64_t n_bytes = end - begin;
    const int64_t tiles = (end - (begin & ~(int64_t)15) + TILE_BYTES - 1) / TILE_BYTES;
    if (!ws.h_total) AGX_CUDA(cudaMallocHost(&ws.h_total, 4 * sizeof(int64_t)));

    auto align = [](int64_t x) { return (x + 255) / 256 * 256; };
    // first stage: tile counts, their scan, total
    const int64_t sz_tc = align(tiles * 4), sz_tb = align(tiles * 8), sz_tmp = align((tiles / SCAN_TILE + 2) * 8);
    int64_t need = sz_tc + sz_tb + sz_tmp + 256;
    auto reserve = [&](int64_t bytes) -> int {
        if (bytes > ws.cap) {
            if (ws.buf) cudaFree(ws.buf);
            ws.buf = nullptr; ws.cap = 0;
            AGX_CUDA(cudaMalloc(&ws.buf, (size_t)bytes));
            ws.cap = bytes;
        }
        return AGX_OK;
    };
    int rc = reserve(need);
    if (rc != AGX_OK) return rc;
    uint8_t *base = reinterpret_cast<uint8_t *>(ws.buf);
    int32_t *tile_count = reinterpret_cast<int32_t *>(base);
    int64_t *tile_base64 = reinterpret_cast<int64_t *>(base + sz_tc);
    int64_t *tmp = reinterpret_cast<int64_t *>(base + sz_tc + sz_tb);
    int64_t *d_total = reinterpret_cast<int64_t *>(base + sz_tc + sz_tb + sz_tmp);

    nl_count_kernel<<<(int)tiles, TILE_THREADS, 0, st>>>(d_img, begin, end, tile_count);
    count_launch();
    AGX_CUDA(cudaGetLastError());
    rc = exclusive_scan(tile_count, tiles, tile_base64, tmp, d_total, st);
    if (rc != AGX_OK) return rc;
    AGX_CUDA(cudaMemcpyAsync(ws.h_total, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    AGX_CUDA(cudaStreamSynchronize(st));
    const int64_t n_nl = ws.h_total[0];
    // does the region end with a newline?  (one byte read unless the caller knows)
    uint8_t last = (uint8_t)last_byte;
    if (last_byte < 0) {
        AGX_CUDA(cudaMemcpyAsync(&last, d_img + end - 1, 1, cudaMemcpyDeviceToHost, st));
        AGX_CUDA(cudaStreamSynchronize(st));
    }
    const int64_t n_lines = n_nl + (last == '\n' ? 0 : 1);
    if (n_nl > (int64_t)1 << 31) return fail(AGX_ERANGE, "sw: more than 2^31 lines");

    // second stage lives in a separate allocation so the first one stays valid
    const int64_t sz_np = align(std::max<int64_t>(n_nl, 1) * 8), sz_cpl = align(n_lines * 4), sz_cb = align(n_lines * 8),
                  sz_tmp2 = align((n_lines / SCAN_TILE + 2) * 8);
    // chunk table size is known after the chunk scan; lines <= cap bytes give exactly n_lines chunks, and a
    // split line adds at most len/cap more: bound by n_lines + n_bytes / cap
    const int64_t chunk_bound = std::min<int64_t>(max_chunks, n_lines + n_bytes / cap + 1);
    // one spare entry in front of each table: a caller may put the chunk that precedes the region there
    const int64_t sz_off = align((chunk_bound + 1) * 8), sz_len = align((chunk_bound + 1) * 4);
    const int64_t need2 = sz_np + sz_cpl + sz_cb + sz_tmp2 + sz_off + sz_len + 512;
    if (need2 > ws.cap2) {
        if (ws.buf2) cudaFree(ws.buf2);
        ws.buf2 = nullptr; ws.cap2 = 0;
        AGX_CUDA(cudaMalloc(&ws.buf2, (size_t)need2));
        ws.cap2 = need2;
    }
    uint8_t *b2 = reinterpret_cast<uint8_t *>(ws.buf2);
    int64_t *nl_pos = reinterpret_cast<int64_t *>(b2);
    int32_t *cpl = reinterpret_cast<int32_t *>(b2 + sz_np);
    int64_t *chunk_base = reinterpret_cast<int64_t *>(b2 + sz_np + sz_cpl);
    int64_t *tmp2 = reinterpret_cast<int64_t *>(b2 + sz_np + sz_cpl + sz_cb);
    int64_t *off = reinterpret_cast<int64_t *>(b2 + sz_np + sz_cpl + sz_cb + sz_tmp2) + 1;
    int32_t *len = reinterpret_cast<int32_t *>(b2 + sz_np + sz_cpl + sz_cb + sz_tmp2 + sz_off) + 1;
    int64_t *d_total2 = reinterpret_cast<int64_t *>(b2 + sz_np + sz_cpl + sz_cb + sz_tmp2 + sz_off + sz_len);

    nl_emit_kernel<<<(int)tiles, TILE_THREADS, 0, st>>>(d_img, begin, end, tile_base64, nl_pos);
    count_launch();
    AGX_CUDA(cudaGetLastError());

    const int lblocks = (int)((n_lines + 255) / 256);
    line_chunks_kernel<<<lblocks, 256, 0, st>>>(nl_pos, n_nl, begin, end, cap, n_lines, cpl);
    count_launch();
    rc = exclusive_scan(cpl, n_lines, chunk_base, tmp2, d_total2, st);
    if (rc != AGX_OK) return rc;
    chunk_emit_kernel<<<lblocks, 256, 0, st>>>(nl_pos, n_nl, begin, end, cap, n_lines, chunk_base, chunk_bound, off, len,
                                               d_total2);
    count_launch();
    AGX_CUDA(cudaGetLastError());
    AGX_CUDA(cudaMemcpyAsync(ws.h_total + 1, d_total2, 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    AGX_CUDA(cudaStreamSynchronize(st));
    *n_chunks_out = std::min<int64_t>(ws.h_total[1], chunk_bound);
    if (*n_chunks_out > 0) {
        if (last_off) *last_off = ws.h_total[2];
        if (last_len) *last_len = (int32_t)ws.h_total[3];
    }
    *d_off = off;
    *d_len = len;
    return AGX_OK;
}

}  // namespace agx
