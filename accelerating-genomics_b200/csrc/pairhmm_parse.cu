// pairhmm_parse.cu -- device-side parser for the pairHMM/test_set text format.
//
// The reference walks its input with fgets() (pairHMM/antidiagsPairHMM.c:375-441): a batch is a header
// line "<num_read> <num_haplotypes>", num_read read lines of five whitespace-separated fields (bases,
// base / insertion / deletion / gap-continuation qualities; read length = (strlen(line) - 4) / 5, :418)
// and num_haplotypes haplotype lines.  This file builds, ON THE GPU and from the raw file image, every
// index array the PairHMM kernels consume, so a driver uploads the image once and never touches it:
//
//   text_newline_index   (sw_parse.cu) position of every '\n'                               HBM-bound
//   hmm_header_flag      which lines look like a batch header (digits, blanks, digits)
//   scan + hmm_headers   header lines in order, their two counts
//   hmm_check_chain      header b + 1 + reads + haplotypes == header b+1 for every b ?     (regular file)
//   hmm_walk             one thread follows the reference's own walk when the chain check fails
//   scans                first read / first haplotype / first output of every batch
//   hmm_fill_lines       one warp per line: field offsets, lengths, batch ids, output offsets, validation
#include <algorithm>

#include "common.cuh"

namespace agx {

namespace {

constexpr int HMM_LINE_MAX = 5000;   // fgets(line, MAX_READ_LEN*5+1 = 5001): longer lines would be split (:353)

// info block (int64 words) shared with the host
enum {
    PI_HEADERS = 0,      // number of header lines found (scan total)
    PI_IRREGULAR = 1,    // the header chain does not tile the file: fall back to the serial walk
    PI_INCOMPLETE = 2,   // the file ends inside the last batch; 2: the reference says "Error reading haplotypes."
                         // (its haplotype cursor runs ahead of the read cursor, :388-396), 1: "Error reading reads."
                         // (only possible for a batch without haplotypes, :411-414)
    PI_BATCHES = 3,      // complete batches
    PI_READS = 4,
    PI_HAPS = 5,
    PI_OUT = 6,
    PI_ERROR = 7,        // validation: 1 read length, 2 haplotype length, 3 field outside line, 4 long line, 5 too many pairs
    PI_ERROR_AT = 8,     // line of the first validation error
    PI_NEXT_BEGIN = 9,   // byte offset of the header of the dropped (incomplete) last batch, else the region end
    PI_LAST_NR = 10,     // header counts of the last complete batch: what a header with fewer than two integers
    PI_LAST_NH = 11,     // right after it inherits
    PI_WORDS = 12
};

__device__ __forceinline__ bool is_blank(uint32_t c) { return c == ' ' || c == '\t'; }

struct LineSpan { int64_t s, e; };   // [s, e) without the '\n'
struct Bounds { int64_t b, e; };     // the parsed region of the image: bytes [b, e), b is a line start
__device__ __forceinline__ LineSpan line_span(const int64_t *nl_pos, int64_t n_nl, Bounds bd, int64_t k)
{
    LineSpan sp;
    sp.s = k == 0 ? bd.b : nl_pos[k - 1] + 1;
    sp.e = k < n_nl ? nl_pos[k] : bd.e;
    return sp;
}

// sscanf(line, "%d %d", &a, &b) with a = b = 0 beforehand (antidiagsPairHMM.c:375-376)
// sscanf(line, "%d %d", &num_read, &num_haplotypes) (antidiagsPairHMM.c:378): a field that does not parse
// leaves its variable as it was -- the reference declares both once, outside the batch loop (:345-346), so a
// header that holds fewer than two integers keeps the previous batch's count(s).  a / b come in holding those.
__device__ void scan_two_ints(const uint8_t *img, LineSpan sp, int32_t &a, int32_t &b)
{
    int64_t p = sp.s;
    for (int f = 0; f < 2; ++f) {
        while (p < sp.e && (is_blank(img[p]) || img[p] == '\r' || img[p] == '\v' || img[p] == '\f')) ++p;
        bool neg = false;
        if (p < sp.e && (img[p] == '+' || img[p] == '-')) { neg = img[p] == '-'; ++p; }
        if (!(p < sp.e && img[p] >= '0' && img[p] <= '9')) return;
        int64_t v = 0;
        while (p < sp.e && img[p] >= '0' && img[p] <= '9') { v = v * 10 + (img[p] - '0'); if (v > 0x7fffffff) v = 0x7fffffff; ++p; }
        (f == 0 ? a : b) = (int32_t)(neg ? -v : v);
    }
}

// strict shape of a header in a well-formed file: digits, blanks, digits, nothing else
__global__ void __launch_bounds__(256)
hmm_header_flag_kernel(const uint8_t *__restrict__ img, const int64_t *__restrict__ nl_pos, int64_t n_nl, Bounds end,
                       int64_t n_lines, int32_t *__restrict__ flag)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines) return;
    const LineSpan sp = line_span(nl_pos, n_nl, end, k);
    int32_t f = 0;
    if (sp.e - sp.s >= 3 && sp.e - sp.s <= 20) {
        int64_t p = sp.s;
        int d0 = 0, bl = 0, d1 = 0;
        while (p < sp.e && img[p] >= '0' && img[p] <= '9') { ++p; ++d0; }
        while (p < sp.e && is_blank(img[p])) { ++p; ++bl; }
        while (p < sp.e && img[p] >= '0' && img[p] <= '9') { ++p; ++d1; }
        f = (p == sp.e && d0 >= 1 && d0 <= 9 && bl >= 1 && d1 >= 1 && d1 <= 9) ? 1 : 0;
    }
    flag[k] = f;
}

__global__ void __launch_bounds__(256)
hmm_headers_kernel(const uint8_t *__restrict__ img, const int64_t *__restrict__ nl_pos, int64_t n_nl, Bounds end,
                   int64_t n_lines, const int32_t *__restrict__ flag, const int64_t *__restrict__ rank,
                   const uint64_t *__restrict__ walked, int32_t *__restrict__ hdr_line, int32_t *__restrict__ nr, int32_t *__restrict__ nh,
                   int32_t *__restrict__ npairs, int64_t *__restrict__ info)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines || !flag[k]) return;
    const int64_t b = rank[k];
    int32_t a = 0, c = 0;
    if (walked) {
        // the serial walk already resolved this header (counts carried over from the batch before included)
        a = (int32_t)(uint32_t)walked[k];
        c = (int32_t)(uint32_t)(walked[k] >> 32);
    } else {
        // strict-shape headers always hold two integers: nothing to carry
        scan_two_ints(img, line_span(nl_pos, n_nl, end, k), a, c);
    }
    if (a < 0) a = 0;          // for (i = 0; i < num_read; i++) runs zero times
    if (c < 0) c = 0;
    hdr_line[b] = (int32_t)k;
    nr[b] = a;
    nh[b] = c;
    const int64_t prod = (int64_t)a * c;
    if (prod > 0x7fffffff) { info[PI_ERROR] = 5; info[PI_ERROR_AT] = k; }
    npairs[b] = (int32_t)(prod > 0x7fffffff ? 0 : prod);
}

// every header must start where the previous batch ends; the last batch may run past the end of the file
__global__ void __launch_bounds__(256)
hmm_check_chain_kernel(const int32_t *__restrict__ hdr_line, const int32_t *__restrict__ nr,
                       const int32_t *__restrict__ nh, const int64_t *info_in, int64_t n_lines, int64_t *info)
{
    const int64_t H = info_in[PI_HEADERS];
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0 && (H == 0 ? n_lines > 0 : hdr_line[0] != 0)) info[PI_IRREGULAR] = 1;
    if (b >= H) return;
    const int64_t next = (int64_t)hdr_line[b] + 1 + nr[b] + nh[b];
    if (b + 1 < H) {
        if (next != hdr_line[b + 1]) info[PI_IRREGULAR] = 1;
    } else if (next > n_lines) {
        info[PI_INCOMPLETE] = nh[b] > 0 ? 2 : 1;
    } else if (next < n_lines) {
        info[PI_IRREGULAR] = 1;      // lines after the last batch that do not look like a header
    }
}

// the reference's own walk, one thread: flags the lines it would read as headers
__global__ void hmm_walk_kernel(const uint8_t *__restrict__ img, const int64_t *__restrict__ nl_pos, int64_t n_nl,
                                Bounds end, int64_t n_lines, int32_t carry_nr, int32_t carry_nh,
                                int32_t *__restrict__ flag, uint64_t *__restrict__ walked, int64_t *__restrict__ info)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int64_t line = 0;
    info[PI_INCOMPLETE] = 0;
    int32_t num_read = carry_nr, num_haplotypes = carry_nh;      // live across batches, as in the reference
    while (line < n_lines) {
        scan_two_ints(img, line_span(nl_pos, n_nl, end, line), num_read, num_haplotypes);
        const int32_t a = num_read < 0 ? 0 : num_read, c = num_haplotypes < 0 ? 0 : num_haplotypes;
        flag[line] = 1;
        walked[line] = (uint64_t)(uint32_t)num_read | ((uint64_t)(uint32_t)num_haplotypes << 32);
        if (line + 1 + a + c > n_lines) { info[PI_INCOMPLETE] = c > 0 ? 2 : 1; break; }
        line += 1 + (int64_t)a + c;
    }
}

__global__ void hmm_totals_kernel(const int64_t *__restrict__ nl_pos, Bounds bd, const int32_t *__restrict__ hdr_line,
                                  const int64_t *info_in, const int64_t *__restrict__ tot_r,
                                  const int64_t *__restrict__ tot_h, const int64_t *__restrict__ tot_o,
                                  const int32_t *__restrict__ nr, const int32_t *__restrict__ nh,
                                  const int32_t *__restrict__ npairs, int64_t *__restrict__ brs,
                                  int64_t *__restrict__ bhs, int64_t *__restrict__ bos, int64_t *info)
{
    // the scans ran over all H headers; an incomplete last batch is dropped (the reference reports it and
    // stops, keeping what it printed for the earlier batches)
    const int64_t H = info_in[PI_HEADERS];
    const int64_t nb = H - (info_in[PI_INCOMPLETE] ? 1 : 0);
    int64_t r = *tot_r, h = *tot_h, o = *tot_o;
    info[PI_NEXT_BEGIN] = bd.e;
    if (info_in[PI_INCOMPLETE] && H > 0) {
        r -= nr[H - 1]; h -= nh[H - 1]; o -= npairs[H - 1];
        const int32_t k = hdr_line[H - 1];
        info[PI_NEXT_BEGIN] = k == 0 ? bd.b : nl_pos[k - 1] + 1;
    }
    info[PI_BATCHES] = nb;
    info[PI_READS] = r;
    info[PI_HAPS] = h;
    info[PI_OUT] = o;
    if (nb > 0) { info[PI_LAST_NR] = nr[nb - 1]; info[PI_LAST_NH] = nh[nb - 1]; }
    if (nb >= 0) { brs[nb] = r; bhs[nb] = h; bos[nb] = o; }
}

// One warp per line.  Header rank (inclusive) - 1 = batch of the line.
__global__ void __launch_bounds__(128)
hmm_fill_lines_kernel(const uint8_t *__restrict__ img, const int64_t *__restrict__ nl_pos, int64_t n_nl, Bounds end,
                      int64_t n_lines, const int32_t *__restrict__ flag, const int64_t *__restrict__ rank,
                      const int32_t *__restrict__ hdr_line, const int32_t *__restrict__ nr,
                      const int32_t *__restrict__ nh, const int64_t *__restrict__ brs,
                      const int64_t *__restrict__ bhs, const int64_t *__restrict__ bos,
                      const int64_t *info_in, int64_t *__restrict__ read_field_off,
                      int32_t *__restrict__ read_len, int32_t *__restrict__ read_batch,
                      int64_t *__restrict__ read_out_off, int64_t *__restrict__ hap_off,
                      int32_t *__restrict__ hap_len, int32_t *__restrict__ batch_pairs, int64_t *info)
{
    const int lane = threadIdx.x & 31;
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n_lines) return;
    const int64_t b = rank[k] + flag[k] - 1;           // headers at or before this line, minus one
    const int64_t nb = info_in[PI_BATCHES];
    if (b < 0 || b >= nb) return;
    const int64_t i = k - hdr_line[b] - 1;
    if (i < 0) {                                       // the header itself
        if (lane == 0) batch_pairs[b] = (int32_t)(bos[b + 1] - bos[b]);
        return;
    }
    const LineSpan sp = line_span(nl_pos, n_nl, end, k);
    const int64_t l = sp.e - sp.s;
    if (l > HMM_LINE_MAX) {
        if (lane == 0) { info[PI_ERROR] = 4; info[PI_ERROR_AT] = k; }
        return;
    }
    if (i >= nr[b]) {
        const int64_t x = bhs[b] + (i - nr[b]);
        if (lane == 0) {
            hap_off[x] = sp.s;
            hap_len[x] = (int32_t)l;
            if (l < 1) { info[PI_ERROR] = 2; info[PI_ERROR_AT] = k; }
        }
        return;
    }
    // a read line: the five fields as sscanf("%s %s %s %s %s") finds them (:101)
    const int64_t r = brs[b] + i;
    int64_t fo[5] = {sp.e, sp.e, sp.e, sp.e, sp.e};
    int found = 0;
    uint32_t carry = 0;
    for (int64_t base = sp.s; base < sp.e && found < 5; base += 32) {
        const int64_t p = base + lane;
        const uint32_t ch = p < sp.e ? img[p] : (uint32_t)' ';
        const uint32_t nonws = __ballot_sync(0xffffffffu, !is_blank(ch));
        uint32_t starts = nonws & ~((nonws << 1) | carry);
        carry = nonws >> 31;
        while (starts && found < 5) {
            fo[found++] = base + (__ffs(starts) - 1);
            starts &= starts - 1;
        }
    }
    if (lane == 0) {
        int64_t len = (l - 4) / 5;                      // :418
        if (len < 0) len = 0;
#pragma unroll
        for (int f = 0; f < 5; ++f) {
            read_field_off[5 * r + f] = fo[f];
            if (fo[f] + len > end.e) { info[PI_ERROR] = 3; info[PI_ERROR_AT] = k; }
        }
        read_len[r] = (int32_t)len;
        read_batch[r] = (int32_t)b;
        read_out_off[r] = bos[b] + i * (int64_t)nh[b];
        if (len < 1 || len > 8192) { info[PI_ERROR] = 1; info[PI_ERROR_AT] = k; }
    }
}

}  // namespace

void hmm_parse_workspace_free(HmmParseWorkspace &ws)
{
    sw_parse_workspace_free(ws.idx);
    if (ws.buf) cudaFree(ws.buf);
    if (ws.tables) cudaFree(ws.tables);
    if (ws.arrays) cudaFree(ws.arrays);
    if (ws.h_info) cudaFreeHost(ws.h_info);
    ws = HmmParseWorkspace();
}

// Parses d_img[begin, bytes) (begin = the start of a header line) into the arrays of `out` (all device
// pointers into the workspace; offsets are offsets into d_img).  A last batch that runs past `bytes` is
// dropped and reported through out->incomplete / out->next_begin, so a caller that holds only a prefix of the
// file can parse it region by region.  Synchronises `st` four times.
int hmm_parse_device(HmmParseWorkspace &ws, const uint8_t *d_img, int64_t begin, int64_t bytes, int last_byte,
                     HmmParsed *out, cudaStream_t st, int32_t carry_nr, int32_t carry_nh)
{
    *out = HmmParsed();
    out->next_begin = bytes;
    if (bytes <= begin) return AGX_OK;
    const Bounds bd{begin, bytes};
    if (!ws.h_info) AGX_CUDA(cudaMallocHost(&ws.h_info, PI_WORDS * sizeof(int64_t)));
    auto align = [](int64_t x) { return (x + 255) / 256 * 256; };

    int64_t *nl_pos = nullptr;
    int64_t n_nl = 0;
    int rc = text_newline_index(ws.idx, d_img, begin, bytes, 0, &nl_pos, &n_nl, nullptr, st);
    if (rc != AGX_OK) return rc;
    const int64_t n_lines = n_nl + (last_byte == '\n' ? 0 : 1);
    if (n_lines == 0) return AGX_OK;

    // per-line scratch: flag, rank; scan temporaries; info block
    const int64_t scan_tmp = device_scan_tmp_elems(n_lines);
    const int64_t sz_flag = align(n_lines * 4), sz_rank = align(n_lines * 8), sz_tmp = align(scan_tmp * 8);
    const int64_t need = sz_flag + 2 * sz_rank + 4 * sz_tmp + 2 * 256;       // + the walk's per-line counts
    if (need > ws.cap) {
        if (ws.buf) cudaFree(ws.buf);
        ws.buf = nullptr; ws.cap = 0;
        AGX_CUDA(cudaMalloc(&ws.buf, (size_t)need));
        ws.cap = need;
    }
    uint8_t *wb = reinterpret_cast<uint8_t *>(ws.buf);
    int32_t *flag = reinterpret_cast<int32_t *>(wb);
    int64_t *rank = reinterpret_cast<int64_t *>(wb + sz_flag);
    int64_t *tmp = reinterpret_cast<int64_t *>(wb + sz_flag + sz_rank);
    int64_t *info = reinterpret_cast<int64_t *>(wb + sz_flag + sz_rank + 4 * sz_tmp);
    uint64_t *walked = reinterpret_cast<uint64_t *>(wb + sz_flag + sz_rank + 4 * sz_tmp + 2 * 256);
    bool use_walk = false;
    int64_t *tot = info + PI_WORDS;                      // three scan totals
    AGX_CUDA(cudaMemsetAsync(info, 0, 256 + 64, st));

    const int lblocks = (int)((n_lines + 255) / 256);
    hmm_header_flag_kernel<<<lblocks, 256, 0, st>>>(d_img, nl_pos, n_nl, bd, n_lines, flag);
    count_launch();
    int64_t H = 0;
    int32_t *hdr_line = nullptr, *nr = nullptr, *nh = nullptr, *npairs = nullptr;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if ((rc = device_exclusive_scan(flag, n_lines, rank, tmp, info + PI_HEADERS, st)) != AGX_OK) return rc;
        AGX_CUDA(cudaMemcpyAsync(ws.h_info, info, PI_WORDS * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        AGX_CUDA(cudaStreamSynchronize(st));
        H = ws.h_info[PI_HEADERS];
        // per-batch tables: hdr_line, nr, nh, npairs (int32), brs, bhs, bos (int64, H+1 each)
        const int64_t sz_i32 = align((H + 1) * 4), sz_i64 = align((H + 2) * 8);
        const int64_t need_t = 5 * sz_i32 + 3 * sz_i64;
        if (need_t > ws.cap_tables) {
            if (ws.tables) cudaFree(ws.tables);
            ws.tables = nullptr; ws.cap_tables = 0;
            AGX_CUDA(cudaMalloc(&ws.tables, (size_t)need_t));
            ws.cap_tables = need_t;
        }
        uint8_t *tb = reinterpret_cast<uint8_t *>(ws.tables);
        hdr_line = reinterpret_cast<int32_t *>(tb);
        nr = reinterpret_cast<int32_t *>(tb + sz_i32);
        nh = reinterpret_cast<int32_t *>(tb + 2 * sz_i32);
        npairs = reinterpret_cast<int32_t *>(tb + 3 * sz_i32);
        out->batch_pairs = reinterpret_cast<int32_t *>(tb + 4 * sz_i32);
        out->batch_read_start = reinterpret_cast<int64_t *>(tb + 5 * sz_i32);
        out->batch_hap_start = reinterpret_cast<int64_t *>(tb + 5 * sz_i32 + sz_i64);
        out->batch_out_start = reinterpret_cast<int64_t *>(tb + 5 * sz_i32 + 2 * sz_i64);
        if (H > 0) {
            hmm_headers_kernel<<<lblocks, 256, 0, st>>>(d_img, nl_pos, n_nl, bd, n_lines, flag, rank,
                                                        use_walk ? walked : nullptr, hdr_line, nr, nh, npairs, info);
            count_launch();
        }
        if (attempt == 0) {
            hmm_check_chain_kernel<<<(int)((std::max<int64_t>(H, 1) + 255) / 256), 256, 0, st>>>(hdr_line, nr, nh, info,
                                                                                                  n_lines, info);
            count_launch();
            AGX_CUDA(cudaMemcpyAsync(ws.h_info, info, PI_WORDS * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
            AGX_CUDA(cudaStreamSynchronize(st));
            if (!ws.h_info[PI_IRREGULAR]) break;
            // not a tiling of header-shaped lines: follow the reference's walk, then redo scan + headers
            AGX_CUDA(cudaMemsetAsync(flag, 0, (size_t)n_lines * sizeof(int32_t), st));
            AGX_CUDA(cudaMemsetAsync(info, 0, PI_WORDS * sizeof(int64_t), st));
            hmm_walk_kernel<<<1, 32, 0, st>>>(d_img, nl_pos, n_nl, bd, n_lines, carry_nr, carry_nh, flag, walked, info);
            count_launch();
            use_walk = true;
        }
    }
    AGX_CUDA(cudaGetLastError());
    if (H == 0) return AGX_OK;
    if ((rc = device_exclusive_scan(nr, H, out->batch_read_start, tmp + scan_tmp, tot + 0, st)) != AGX_OK) return rc;
    if ((rc = device_exclusive_scan(nh, H, out->batch_hap_start, tmp + 2 * scan_tmp, tot + 1, st)) != AGX_OK) return rc;
    if ((rc = device_exclusive_scan(npairs, H, out->batch_out_start, tmp + 3 * scan_tmp, tot + 2, st)) != AGX_OK) return rc;
    hmm_totals_kernel<<<1, 1, 0, st>>>(nl_pos, bd, hdr_line, info, tot + 0, tot + 1, tot + 2, nr, nh, npairs, out->batch_read_start,
                                       out->batch_hap_start, out->batch_out_start, info);
    count_launch();
    AGX_CUDA(cudaMemcpyAsync(ws.h_info, info, PI_WORDS * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    AGX_CUDA(cudaStreamSynchronize(st));
    if (ws.h_info[PI_ERROR] == 5) return fail(AGX_ERANGE, "pairhmm: a batch holds more than 2^31 pairs");
    out->incomplete = (int32_t)ws.h_info[PI_INCOMPLETE];
    out->next_begin = ws.h_info[PI_NEXT_BEGIN];
    out->last_nr = carry_nr;
    out->last_nh = carry_nh;
    if (ws.h_info[PI_BATCHES] > 0) { out->last_nr = (int32_t)ws.h_info[PI_LAST_NR]; out->last_nh = (int32_t)ws.h_info[PI_LAST_NH]; }
    out->n_batches = ws.h_info[PI_BATCHES];
    out->n_reads = ws.h_info[PI_READS];
    out->n_haps = ws.h_info[PI_HAPS];
    out->n_out = ws.h_info[PI_OUT];
    if (out->n_batches <= 0) { out->n_batches = 0; return AGX_OK; }

    // per-read / per-haplotype arrays
    const int64_t nrd = std::max<int64_t>(out->n_reads, 1), nhp = std::max<int64_t>(out->n_haps, 1);
    const int64_t sz_rfo = align(nrd * 5 * 8), sz_roo = align(nrd * 8), sz_ho = align(nhp * 8), sz_rl = align(nrd * 4),
                  sz_rb = align(nrd * 4), sz_hl = align(nhp * 4);
    const int64_t need_a = sz_rfo + sz_roo + sz_ho + sz_rl + sz_rb + sz_hl;
    if (need_a > ws.cap_arrays) {
        if (ws.arrays) cudaFree(ws.arrays);
        ws.arrays = nullptr; ws.cap_arrays = 0;
        AGX_CUDA(cudaMalloc(&ws.arrays, (size_t)need_a));
        ws.cap_arrays = need_a;
    }
    uint8_t *ab = reinterpret_cast<uint8_t *>(ws.arrays);
    out->read_field_off = reinterpret_cast<int64_t *>(ab);
    out->read_out_off = reinterpret_cast<int64_t *>(ab + sz_rfo);
    out->hap_off = reinterpret_cast<int64_t *>(ab + sz_rfo + sz_roo);
    out->read_len = reinterpret_cast<int32_t *>(ab + sz_rfo + sz_roo + sz_ho);
    out->read_batch = reinterpret_cast<int32_t *>(ab + sz_rfo + sz_roo + sz_ho + sz_rl);
    out->hap_len = reinterpret_cast<int32_t *>(ab + sz_rfo + sz_roo + sz_ho + sz_rl + sz_rb);

    const int64_t fblocks = (n_lines * 32 + 127) / 128;
    hmm_fill_lines_kernel<<<(int)fblocks, 128, 0, st>>>(d_img, nl_pos, n_nl, bd, n_lines, flag, rank, hdr_line, nr, nh,
                                                        out->batch_read_start, out->batch_hap_start,
                                                        out->batch_out_start, info, out->read_field_off, out->read_len,
                                                        out->read_batch, out->read_out_off, out->hap_off, out->hap_len,
                                                        out->batch_pairs, info);
    count_launch();
    AGX_CUDA(cudaGetLastError());
    AGX_CUDA(cudaMemcpyAsync(ws.h_info, info, PI_WORDS * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    AGX_CUDA(cudaStreamSynchronize(st));
    const int64_t at = ws.h_info[PI_ERROR_AT] + 1;
    switch (ws.h_info[PI_ERROR]) {
    case 0: break;
    case 1: return fail(AGX_ERANGE, "pairhmm: read on line " + std::to_string(at) + " has a length outside [1, 8192]");
    case 2: return fail(AGX_ERANGE, "pairhmm: haplotype on line " + std::to_string(at) + " is empty");
    case 3: return fail(AGX_EINVAL, "pairhmm: read on line " + std::to_string(at) + " has fewer than five full fields");
    case 4: return fail(AGX_ERANGE, "pairhmm: line " + std::to_string(at) + " is longer than the reference's 5000-byte line buffer");
    default: return fail(AGX_EINVAL, "pairhmm: malformed input");
    }
    return AGX_OK;
}

}  // namespace agx
