// common.cuh -- shared declarations for libagx (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/agx.h"

namespace agx {

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define AGX_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return ::agx::fail(AGX_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- optional per-kernel timing (agx_set_profiling): CUDA events around the dominant kernels ----
struct ProfSpan {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool armed = false;
    void begin(cudaStream_t st);
    void end(cudaStream_t st);
    double ms();          // synchronises e1; < 0 when nothing was recorded
    void destroy();
};
extern std::atomic<int> g_profiling;
enum { AGX_PROF_SW_DUO = 0, AGX_PROF_SW_WAVE = 1, AGX_PROF_HMM_STREAM = 2, AGX_PROF_HMM_FP64 = 3,
       AGX_PROF_SW_CLASSIFY = 4, AGX_PROF_HMM_CLASSIFY = 5, AGX_PROF_SW_LONG = 6, AGX_PROF_SW_ALIGN_DP = 7,
       AGX_PROF_SW_ALIGN_WALK = 8, AGX_PROF_COUNT = 9 };

// ---- Smith-Waterman -----------------------------------------------------------------------
struct SwScoring {
    int32_t match, mismatch, gap_open, gap_extend;
};

// Number of length classes handled by the packed s16x2 inter-task kernel; the extra class
// (index SW_N_DUO_CLASSES) is the generic s32 wavefront kernel.
constexpr int SW_N_DUO_CLASSES = 11;
// + the generic one-warp-per-pair wavefront class + the whole-GPU long-alignment class
constexpr int SW_N_CLASSES = SW_N_DUO_CLASSES + 2;
// a pair goes to the whole-GPU kernel (sw_long.cu) when it has at least this many cells and its
// shorter side does not fit the inter-task kernel
constexpr int64_t SW_LONG_CELLS_DEFAULT = (int64_t)1 << 28;
int64_t sw_long_cells();   // SW_LONG_CELLS_DEFAULT unless the environment sets AGX_SW_LONG_CELLS

// scratch of the long-alignment kernel
struct SwLongWorkspace {
    int32_t *buf = nullptr;   // boundary column (2 * rows) + per-stripe progress counters + best
    int64_t cap = 0;          // in int32 elements
    uint8_t *seq = nullptr;   // device copies of the two sequences (host entry point)
    int64_t cap_seq = 0;
    ProfSpan prof;            // the long-alignment kernel of the last host-entry call on this GPU
};
int sw_long_device(SwLongWorkspace &ws, const uint8_t *d_a, int64_t la, const uint8_t *d_b, int64_t lb,
                   SwScoring sc, int32_t *d_best, cudaStream_t st);
// end_col_out / end_row_out (both or neither): 0-based END CELL, the cell the reference's running maximum comes from
// (-1 -1 when the score is 0); col_is_sx: the reference's ix walks `a` (else `b`)
int sw_long_host_multi(int n_dev, const int *dev, cudaStream_t *st, SwLongWorkspace **ws, const uint8_t *a,
                       int64_t la, const uint8_t *b, int64_t lb, SwScoring sc, int32_t *score_out,
                       int32_t *end_col_out = nullptr, int32_t *end_row_out = nullptr, int col_is_sx = 0);
void sw_long_workspace_free(SwLongWorkspace &ws);

// ---- Smith-Waterman alignment: end cell, start cell, CIGAR (sw_align.cuh) -------------------------------------
// One per pair, written by the DP kernels (MODE 2), read by the traceback walk.
struct SwWalkRec {
    int64_t tb_off;            // byte offset of the H-byte matrix (of the pair's duo / of the pair) in the scratch
    int32_t r_end, c_end;      // cell the walk starts from, kernel row / column
    int32_t row_off, col_off;  // kernel row / column of symbol 0 of the row / column sequence
    int32_t rstride;           // rows of one strip (duo layout) / rows of the pair (wavefront layout)
    int16_t cls;               // duo class, -1 = wavefront layout
    uint8_t half;              // 16-bit half of the duo
    uint8_t flags;
};
static_assert(sizeof(SwWalkRec) == 32, "SwWalkRec");
enum { SW_WK_A_IS_X = 1,       // the column sequence is line 1
       SW_WK_NL_END = 2,       // the alignment ends on the newline symbols; the walk starts one cell before
       SW_WK_NONE = 4,         // score 0: no alignment
       SW_WK_RAW = 8,          // the kernel saw raw lines (newline symbols are ordinary rows / columns)
       SW_WK_TRIVIAL = 16 };   // a line holds nothing but its newline: score and the one-cell alignment are known

struct SwAlignWorkspace {
    int32_t *order = nullptr;        // [SW_N_CLASSES][n_pairs]
    int32_t *cap32 = nullptr;        // [n_pairs] most CIGAR runs a pair can have; later its run count
    int64_t *tmp_off = nullptr;      // [n_pairs + 1] exclusive scan of cap32
    int64_t *cig_off = nullptr;      // [n_pairs + 1] exclusive scan of the run counts
    int32_t *gen_units = nullptr;    // [n_pairs] 256-byte units of each listed wavefront pair's matrix; later the walk order
    int32_t *walk_bins = nullptr;    // score-size histogram + cursors of that order
    int64_t *gen_off = nullptr;      // [n_pairs + 1] their scan (in units, turned into bytes in place)
    int64_t *scan_tmp = nullptr;
    int64_t *d_total = nullptr;      // 4 totals
    int64_t *h_total = nullptr;      // pinned
    SwWalkRec *wk = nullptr;         // [n_pairs]
    int64_t cap_pairs = 0;
    int32_t *counters = nullptr, *h_counters = nullptr;
    uint8_t *tb = nullptr;           // H-byte matrices of the duo classes
    int64_t cap_tb = 0;
    uint8_t *tb_gen = nullptr;       // ... of the wavefront pairs (sized after the duo kernels bounced theirs)
    int64_t cap_tb_gen = 0;
    uint32_t *tmp_ops = nullptr;     // runs as the walk finds them (end -> start)
    int64_t cap_tmp = 0;
    int32_t *wave_scratch = nullptr;
    int64_t cap_wave = 0;
    ProfSpan prof_dp, prof_walk;
    double dp_ms_sum = -1.0, walk_ms_sum = -1.0;   // over the chunks of the last host call (profiling on)
};
// mode 1: scores + end cells; mode 2: + start cells, run counts and (in ws.tmp_ops, reversed) the CIGAR runs.
// d_ends [2n] (index in line 1, index in line 2); d_coords [4n] a_start a_end b_start b_end (mode 2).
// tb_budget: bytes the H-byte matrices of this call may take (AGX_ENOMEM beyond it: the caller cuts smaller chunks).
int sw_align_run_device(SwAlignWorkspace &ws, const uint8_t *d_seqs, const int64_t *d_off, const int32_t *d_len,
                        int64_t n_pairs, SwScoring sc, int mode, int64_t tb_budget, int32_t *d_scores, int32_t *d_ends,
                        int32_t *d_coords, int64_t *cigar_total, cudaStream_t st);
// after a mode-2 run: runs of pair p go to d_cigar[ws.cig_off[p] ...), start -> end
int sw_align_gather_device(SwAlignWorkspace &ws, int64_t n_pairs, uint32_t *d_cigar, cudaStream_t st);
void sw_align_workspace_free(SwAlignWorkspace &ws);
// bytes per row of the H-byte matrix one pair can take (host-side chunking)
int64_t sw_align_tb_row_bytes(int32_t len_a, int32_t len_b);
// the table behind it, indexed by the shorter length up to *max_len (beyond: 256-byte stripes)
const int32_t *sw_align_tb_row_table(int32_t *max_len);

// Per-call device scratch for the SW path (owned by the device context).
struct SwWorkspace {
    int32_t *order = nullptr;        // [n_pairs]   pair ids grouped by class
    int32_t *pair_class = nullptr;   // [n_pairs]
    int32_t *counters = nullptr;     // [SW_N_CLASSES] histogram, then [SW_N_CLASSES] cursors, then misc
    int32_t *wave_scratch = nullptr; // boundary columns for the wavefront kernel
    int64_t cap_pairs = 0;
    int64_t cap_wave = 0;
    int32_t *h_counters = nullptr;   // pinned mirror of counters
    ProfSpan prof_duo, prof_wave, prof_classify, prof_long;
    SwLongWorkspace lng;
};

// Enqueue the whole SW path for one device-resident batch on `st`.  With prep_st (a high-priority
// stream) the length-class pass runs there, so it is not queued behind DP kernels of another batch; the
// caller then guarantees that `ws` and the inputs are no longer in use by earlier work on `st`.
int sw_run_device(SwWorkspace &ws, const uint8_t *d_seqs, const int64_t *d_off, const int32_t *d_len,
                  int64_t n_pairs, SwScoring sc, int32_t *d_scores, cudaStream_t st, cudaStream_t prep_st = nullptr);
int sw_workspace_reserve(SwWorkspace &ws, int64_t n_pairs);
void sw_workspace_free(SwWorkspace &ws);

// device-side fgets() chunking of a Smith-Waterman file image (sw_parse.cu)
struct SwParseWorkspace {
    void *buf = nullptr;      // tile counts / bases
    int64_t cap = 0;
    void *buf2 = nullptr;     // newline positions, chunks per line, chunk table
    int64_t cap2 = 0;
    int64_t *h_total = nullptr;
};
int sw_parse_device(SwParseWorkspace &ws, const uint8_t *d_img, int64_t begin, int64_t end, int32_t line_buf,
                    int64_t max_chunks, int last_byte, int64_t **d_off, int32_t **d_len, int64_t *n_chunks_out,
                    int64_t *last_off, int32_t *last_len, cudaStream_t st);
void sw_parse_workspace_free(SwParseWorkspace &ws);
// shared text-indexing pieces (sw_parse.cu), also used by the PairHMM file parser
int text_newline_index(SwParseWorkspace &ws, const uint8_t *d_img, int64_t begin, int64_t end, int64_t extra_bytes_per_line,
                       int64_t **d_nl_pos, int64_t *n_nl_out, void **d_extra, cudaStream_t st);
int device_exclusive_scan(const int32_t *in, int64_t n, int64_t *out, int64_t *tmp, int64_t *d_total, cudaStream_t st);
int64_t device_scan_tmp_elems(int64_t n);

// ---- PairHMM ------------------------------------------------------------------------------
struct HmmWorkspace {
    int32_t *order = nullptr;     // [n_reads] read ids grouped by row class
    void *pairs = nullptr;        // [HMM_MAX_K][n_reads] int2: reads paired for the two-reads-per-warp kernel
    int32_t *batch_first = nullptr;   // [n_batches] first read, then [n_batches] one past the last read
    int64_t cap_batches = 0;
    int32_t *counters = nullptr;  // class histogram / cursors / rescue count
    int32_t *h_counters = nullptr;
    int64_t *rescue = nullptr;    // [cap_pairs] flat output indices needing FP64
    int64_t cap_reads = 0;
    int64_t cap_pairs = 0;
    int64_t last_rescue = -1;     // pairs re-run in FP64 by the last call that read the count back
    double *d_lut = nullptr;      // 256-entry Phred+33 -> probability table (host libm pow)
    void *scratch = nullptr;      // boundary rows of the striped kernel
    int64_t cap_scratch = 0;
    void *prep = nullptr;         // haplotype codes + records + FP32 forward sums of the stream kernel
    int64_t cap_prep = 0;
    ProfSpan prof_stream, prof_fp64, prof_classify;
    // the row-class launches of the stream kernel run side by side so that one class's tail wave
    // overlaps the next class's first wave
    cudaStream_t aux[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
};

struct HmmBatchView {
    const uint8_t *buf;
    const int64_t *read_field_off;  // [5*n_reads]
    const int32_t *read_len;        // [n_reads]
    const int32_t *read_batch;      // [n_reads]
    int64_t n_reads;
    const int64_t *hap_off;
    const int32_t *hap_len;
    int64_t n_haps;
    const int64_t *batch_hap_start; // [n_batches+1]
    int64_t n_batches;
};

// device-side parser of the pairHMM/test_set text format (pairhmm_parse.cu)
struct HmmParseWorkspace {
    SwParseWorkspace idx;     // newline index
    void *buf = nullptr;      // per-line flags / ranks, scan temporaries, info block
    int64_t cap = 0;
    void *tables = nullptr;   // per-batch tables
    int64_t cap_tables = 0;
    void *arrays = nullptr;   // per-read / per-haplotype arrays
    int64_t cap_arrays = 0;
    int64_t *h_info = nullptr;
};
struct HmmParsed {            // device pointers into the workspace
    int64_t n_batches = 0, n_reads = 0, n_haps = 0, n_out = 0;
    int32_t incomplete = 0;   // the file ends inside a last batch (dropped): 2 = "Error reading haplotypes.", 1 = "... reads."
    int64_t next_begin = 0;   // where parsing resumes: the dropped batch's header, else the end of the region
    int32_t last_nr = 0, last_nh = 0;   // header counts of the last complete batch (what a following header inherits)
    int64_t *read_field_off = nullptr, *read_out_off = nullptr, *hap_off = nullptr;
    int32_t *read_len = nullptr, *read_batch = nullptr, *hap_len = nullptr, *batch_pairs = nullptr;
    int64_t *batch_read_start = nullptr, *batch_hap_start = nullptr, *batch_out_start = nullptr;
};
// carry_nr / carry_nh: the counts of the batch parsed before `begin` (0, 0 at the start of a file): a header with
// fewer than two integers keeps them, as the reference's sscanf() does (antidiagsPairHMM.c:345-346, :378)
int hmm_parse_device(HmmParseWorkspace &ws, const uint8_t *d_img, int64_t begin, int64_t bytes, int last_byte,
                     HmmParsed *out, cudaStream_t st, int32_t carry_nr = 0, int32_t carry_nh = 0);
void hmm_parse_workspace_free(HmmParseWorkspace &ws);

int hmm_workspace_reserve(HmmWorkspace &ws, int64_t n_reads, int64_t n_pairs, int64_t n_batches);
void hmm_workspace_free(HmmWorkspace &ws);
// d_read_out_off[r] = index in d_out of (read r, first haplotype of its batch); n_pairs = total outputs.
// rescue: 0 = leave pairs FP32 cannot be trusted with as NaN, 1 = re-run them in FP64 (one more stream
// synchronisation to size that launch), 2 = the same without the synchronisation (the FP64 kernel reads the
// count on the device; it is launched even when there is nothing to do).
// gatk_mode: 0 the reference's priors, bit 0 mismatch prior Qr/3, bit 1 base-quality floor 6 (see agx.h)
int hmm_run_device(HmmWorkspace &ws, const HmmBatchView &v, int64_t buf_bytes, const int64_t *d_read_out_off,
                   int64_t n_pairs, int gatk_mode, bool force_fp64, int rescue, double *d_out,
                   cudaStream_t st, cudaStream_t prep_st = nullptr);

}  // namespace agx
