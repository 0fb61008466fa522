/*
 * agx.h -- C ABI of libagx.so: the B200 (sm_100a) implementation of the two dynamic-programming
 * hot paths of AnteMarusic/Accelerating-Genomics.
 *
 * The reference has NO function-level FFI: each program is one main().  The
 * entry points below are therefore the ones BASELINE.json's north_star prescribes
 * (sw_score_batch / pairhmm_forward_batch); each cites the reference code it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *   - plain pointers and sizes only; caller owns every host buffer; the library owns device memory,
 *     streams and staging buffers; outputs are complete when a call returns (host variants) or are
 *     ordered on `stream` (device variants).
 *   - return 0 on success, a negative AGX_E* code on failure; agx_last_error() describes the last
 *     failure of the calling thread.  The library never calls exit() and has NO CPU fallback: without
 *     a usable CUDA device every compute entry point fails with AGX_ENODEVICE.
 *   - sequences are RAW BYTES exactly as the reference program sees them.  For Smith-Waterman that
 *     includes the trailing '\n' of each line, which the reference scores as a symbol
 *     (antidiagonalSmithWaterman.c:229-244: strlen() counts it; :332 compares bytes).
 *   - host entry points are synchronous and not re-entrant (the reference is single-threaded);
 *     internally they use one host thread + one stream per configured GPU.
 */
#ifndef AGX_H
#define AGX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGX_OK          0
#define AGX_EINVAL     -1   /* bad argument                                   */
#define AGX_ENODEVICE  -2   /* no usable CUDA device / library not initialised */
#define AGX_ECUDA      -3   /* a CUDA runtime call or kernel failed            */
#define AGX_ENOMEM     -4   /* host or device allocation failed                */
#define AGX_ERANGE     -5   /* input outside the supported range (see call)    */

/* ------------------------------------------------------------------ runtime */

/* Bind the library to GPUs.  n_gpus <= 0 uses every visible device, otherwise devices 0..n_gpus-1.
 * Replaces the hard-coded `int dev = 1; cudaSetDevice(dev)` of smithWaterman.cu:391 /
 * pairHMM.cu:376.  Idempotent; may be called again after agx_shutdown(). */
int agx_init(int32_t n_gpus);

/* Same, with an explicit device list (one process per GPU under torchrun passes {LOCAL_RANK}). */
int agx_init_devices(const int32_t *device_ids, int32_t n_devices);

/* Number of GPUs the library is bound to (0 before agx_init). */
int32_t agx_device_count(void);

/* CUDA ordinal and name of the index-th bound GPU (what smithWaterman.cu:393 / hipvers.cpp:389 print as
 * "[main] Using Device %d: %s"); -1 / "" when index is out of range.  The string lives until agx_shutdown(). */
int32_t agx_device_ordinal(int32_t index);
const char *agx_device_name(int32_t index);

/* Release every device/host resource.  Safe to call when not initialised. */
void agx_shutdown(void);

/* Message for the last error on the calling thread ("" if none).  Never NULL. */
const char *agx_last_error(void);

/* Library version string, e.g. "agx 0.1 sm_100a". */
const char *agx_version(void);

/* Kernel launches issued by this process since the last agx_reset_launch_count() (all devices).
 * bench.py reports this as "gpu_launches". */
int64_t agx_launch_count(void);
void agx_reset_launch_count(void);

/* Optional per-kernel timing for bench.py's roofline line.  With profiling on, the library brackets
 * its dominant kernels with CUDA events on the stream they are launched on; agx_profile_ms()
 * returns the duration (ms) of the LAST recorded span on `device`, or a negative number if that
 * span never ran.  which: 0 SW inter-task (duo) kernels, 1 SW wavefront kernel, 2 PairHMM FP32
 * stream kernels, 3 PairHMM FP64 kernel, 4 SW classify kernel, 5 PairHMM classify kernel,
 * 6 SW whole-GPU long-alignment kernel(s), 7 DP kernels of the last sw_ends_* / sw_align_* call (summed over its chunks), 8 its traceback walks. */
int agx_set_profiling(int32_t on);
double agx_profile_ms(int32_t device, int32_t which);

/* ------------------------------------------------------- Smith-Waterman (score only) */

/* Scores n_pairs independent pairs; pair p is (a[p], a_len[p]) vs (b[p], b_len[p]).
 * Replaces the per-pair DP loop of antidiagonalSmithWaterman.c:246-348 (boundary init :290-306,
 * P/Q/D recurrence :309-335, running max :335) and its rolling store m_get/m_set :96-184.
 * Scoring uses the reference's sign convention: match > 0 > mismatch, gap_open <= 0,
 * gap_extend < 0, the first base of a gap costs gap_open + gap_extend (:313, :321).  The
 * reference's constants are (1, -1, -3, -1) (:40-43).  Other sign combinations -> AGX_ERANGE.
 * scores_out[p] is bit-exact with the reference's `Score: %d` (:348).
 * Pairs are sharded across the configured GPUs by cell count; no collective is involved.
 * A pair with >= 2^28 cells whose shorter side exceeds 1024 symbols is scored on its own by the
 * intra-task wavefront kernel: its columns are striped over all SMs and, in the host entry points,
 * over all configured GPUs (boundary columns move between neighbouring GPUs as NVLink peer stores).
 * The threshold can be changed with the environment variable AGX_SW_LONG_CELLS. */
int sw_score_batch(const uint8_t *const *a, const int32_t *a_len,
                   const uint8_t *const *b, const int32_t *b_len, int64_t n_pairs,
                   int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                   int32_t *scores_out);

/* Same computation on a flat buffer: sequence 2p is `a` of pair p, 2p+1 is `b` -- the line order of
 * a generator.py file (smithWaterman/generator.py:22-26), so a driver can pass the file image and
 * the (offset, length) of every fgets() chunk without copying.  off[i]/len[i] address seqs. */
int sw_score_batch_flat(const uint8_t *seqs, int64_t seqs_bytes, const int64_t *off,
                        const int32_t *len, int64_t n_pairs,
                        int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                        int32_t *scores_out);

/* Scores a whole generator.py-style FILE IMAGE; the fgets() chunking of
 * antidiagonalSmithWaterman.c:201-227 is reproduced on the GPU (sw_parse.cu), so the caller builds no
 * (offset, length) arrays: line 1 is atoi()'d into the number of LINES to consume (:209), every
 * following sequence is one fgets(line_buf) chunk -- at most line_buf-1 bytes, '\n' kept -- and
 * ceil(lines/2) pairs are scored (:216).  line_buf = 1000 is the reference's MAX_LINE_LENGTH (:44).
 * *n_pairs_out = pairs scored (scores_out[0 .. n)), *header_out = atoi(line 1); when the file ends in
 * the middle of a pair, *dangling_off / *dangling_len locate the first line of that pair (the
 * reference echoes it, :223-227), else *dangling_off = -1.  With several GPUs bound and an image of at least
 * 4 MiB per GPU the image is cut into one byte range per GPU (no collective: every GPU chunks and scores its
 * own range); otherwise the upload is streamed to the first GPU and overlapped with its DP kernels. */
int sw_score_file_image(const uint8_t *image, int64_t image_bytes, int32_t line_buf,
                        int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                        int32_t *scores_out, int64_t scores_cap, int64_t *n_pairs_out,
                        int32_t *header_out, int64_t *dangling_off, int32_t *dangling_len);

/* Device-resident variant for one GPU: every pointer is a device pointer on `device`
 * (a CUDA ordinal that was passed to agx_init*), work is enqueued on `stream`
 * (a cudaStream_t / CUstream cast to void*, NULL = the library's stream for that device).
 * The call sizes its grids from a small device->host histogram read, so it synchronises `stream`
 * once internally before the DP kernels are queued; the kernels themselves are left asynchronous. */
int sw_score_batch_device(int32_t device, const uint8_t *d_seqs, int64_t seqs_bytes,
                          const int64_t *d_off, const int32_t *d_len, int64_t n_pairs,
                          int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                          int32_t *d_scores_out, void *stream);

/* Several device-resident shards in ONE call, one per GPU: the library's multi-GPU dispatcher (one host thread
 * + one stream per GPU, no collective -- pairs are independent) without the host<->device copies, i.e. what
 * sw_score_batch_flat does after its uploads.  Every pointer of shards[k] is a device pointer on
 * shards[k].device (a CUDA ordinal that was passed to agx_init*; each device at most once).  Returns when every
 * shard's scores are complete in its d_scores_out. */
typedef struct {
    int32_t device;
    const uint8_t *d_seqs;
    int64_t seqs_bytes;
    const int64_t *d_off;
    const int32_t *d_len;
    int64_t n_pairs;
    int32_t *d_scores_out;
} agx_sw_shard;
int sw_score_shards_device(const agx_sw_shard *shards, int32_t n_shards,
                           int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend);

/* ------------------------------------------- Smith-Waterman alignment: end cell, start cell, CIGAR */

/* What comes after the score (the reference leaves traceback out, README.md:2).
 * Same recurrence, same raw-byte symbols (a trailing '\n' is a symbol and can be the last aligned column).
 * scores_out[p] equals sw_score_batch_flat's.  ends_out[2p], ends_out[2p+1] = 0-based index in sequence 2p / 2p+1
 * of the LAST aligned symbol: the cell the reference's running maximum comes from
 * (antidiagonalSmithWaterman.c:335 `max = val > max ? val : max`: the first cell with the final value in its
 * visiting order -- anti-diagonals :270-347, inside one with ix ascending, ix walking the shorter line, line 1
 * when both are equally long :229-244).  Both -1 when the score is 0.
 * A whole-GPU pair (>= 2^28 cells with a shorter side above 1024, see sw_score_batch) gets its end cell from the
 * striped long-alignment kernel over every configured GPU, like its score; that needs both lines <= 2^21 - 1 symbols
 * made of at most 7 distinct bytes (else AGX_ERANGE).  Other limits (AGX_ERANGE): match - (gap_open + gap_extend),
 * -mismatch, -(gap_open + gap_extend) <= 127; every other pair: shorter line <= 16000 symbols, the two lengths
 * adding up to less than 2^29. */
int sw_ends_batch_flat(const uint8_t *seqs, int64_t seqs_bytes, const int64_t *off, const int32_t *len,
                       int64_t n_pairs, int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                       int32_t *scores_out, int32_t *ends_out);

/* Full local alignment.  coords_out[4p ..] = a_start, a_end, b_start, b_end (0-based, inclusive, a = sequence 2p,
 * b = sequence 2p+1; all -1 when the score is 0); the runs of pair p are cigar_out[cigar_off_out[p] ..
 * cigar_off_out[p+1]), start -> end, each length << 4 | op with op 0 = M (one symbol of each sequence, equal or
 * not), 1 = I (symbols of a only), 2 = D (symbols of b only) -- BAM's encoding.  cigar_off_out has n_pairs + 1
 * entries.  The path is the one this rule picks from the score matrix, walking back from the end cell: the
 * diagonal when D[i][j] == D[i-1][j-1] + subst, else the shortest gap that explains D[i][j] (symbols of a before
 * symbols of b at equal length), until a cell with D == 0; re-scoring the CIGAR gives exactly scores_out[p].
 * *cigar_total_out = runs of the whole batch; when that exceeds cigar_cap the call returns AGX_ERANGE with
 * scores, coordinates and offsets complete and no runs written beyond the capacity's last whole shard.
 * The score matrices live on the GPU one byte per cell (the batch is cut into chunks that fit device memory);
 * whole-GPU pairs (>= 2^28 cells, see sw_score_batch) are refused here (AGX_ERANGE): use sw_ends_batch_flat. */
int sw_align_batch_flat(const uint8_t *seqs, int64_t seqs_bytes, const int64_t *off, const int32_t *len,
                        int64_t n_pairs, int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                        int32_t *scores_out, int32_t *coords_out, int64_t *cigar_off_out, uint32_t *cigar_out,
                        int64_t cigar_cap, int64_t *cigar_total_out);

/* The same two calls on pointer arrays, pair p = (a[p], a_len[p]) vs (b[p], b_len[p]) as in sw_score_batch. */
int sw_ends_batch(const uint8_t *const *a, const int32_t *a_len, const uint8_t *const *b, const int32_t *b_len,
                  int64_t n_pairs, int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                  int32_t *scores_out, int32_t *ends_out);
int sw_align_batch(const uint8_t *const *a, const int32_t *a_len, const uint8_t *const *b, const int32_t *b_len,
                   int64_t n_pairs, int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                   int32_t *scores_out, int32_t *coords_out, int64_t *cigar_off_out, uint32_t *cigar_out,
                   int64_t cigar_cap, int64_t *cigar_total_out);

/* Device-resident variants for one GPU (every pointer a device pointer on `device`, work ordered on `stream`, see
 * sw_score_batch_device).  The calls read small counters back to size their grids (a few stream synchronisations);
 * sw_align_batch_device keeps the score matrices of the WHOLE batch at once (AGX_ENOMEM when they do not fit the free
 * device memory: cut the batch) and returns *cigar_total_out on the host; whole-GPU pairs are refused (AGX_ERANGE). */
int sw_ends_batch_device(int32_t device, const uint8_t *d_seqs, int64_t seqs_bytes, const int64_t *d_off,
                         const int32_t *d_len, int64_t n_pairs, int32_t match, int32_t mismatch, int32_t gap_open,
                         int32_t gap_extend, int32_t *d_scores_out, int32_t *d_ends_out, void *stream);
int sw_align_batch_device(int32_t device, const uint8_t *d_seqs, int64_t seqs_bytes, const int64_t *d_off,
                          const int32_t *d_len, int64_t n_pairs, int32_t match, int32_t mismatch, int32_t gap_open,
                          int32_t gap_extend, int32_t *d_scores_out, int32_t *d_coords_out, int64_t *d_cigar_off_out,
                          uint32_t *d_cigar_out, int64_t cigar_cap, int64_t *cigar_total_out, void *stream);

/* ------------------------------------------------------------------ PairHMM forward */

/* One batch: every read against every haplotype (the Cartesian loop of
 * antidiagsPairHMM.c:411-469 / pairHMMmatrix.c:207-291).  For read r the five arrays hold
 * read_len[r] bytes: bases and the Phred+33 base / insertion / deletion / gap-continuation
 * qualities (the five fields partition_read() splits, antidiagsPairHMM.c:99-109).
 * Replaces: prior setup partition_read :99-109, p() :111-113 (mismatch prior = Qr, the reference's
 * quirk, reproduced), mm() :115-117, the M/X/Y recurrence pairHMM() :120-241
 * (= pairHMMmatrix.c:41-56) and the final log10 sum :206-212, :242 (= pairHMMmatrix.c:59-66).
 * log10_out[r * n_haps + h] (read-major, the reference's output order :459-461) agrees with the
 * reference's double within 1e-5 relative; pairs whose FP32 forward sum would lose precision are
 * re-run in FP64 on the GPU.  read_len in [1, 8192], hap_len in [1, 2^20] else AGX_ERANGE. */
int pairhmm_forward_batch(int32_t n_reads, const uint8_t *const *bases, const uint8_t *const *q,
                          const uint8_t *const *qi, const uint8_t *const *qd,
                          const uint8_t *const *qg, const int32_t *read_len,
                          int32_t n_haps, const uint8_t *const *haps, const int32_t *hap_len,
                          double *log10_out);

/* Many batches in one call, on a flat buffer (the file image of a pairHMM/test_set style input).
 *   read_field_off[5*r + f]  offset in buf of field f (0 bases, 1 q, 2 qi, 3 qd, 4 qg) of read r
 *   read_len[r]              bases in read r
 *   hap_off[h], hap_len[h]   haplotype h
 *   batch_read_start[b], batch_hap_start[b] (n_batches+1 entries each) delimit batch b
 *   log10_out                sum_b nr_b*nh_b doubles, batches in order, read-major inside a batch
 * Reads are sharded across the configured GPUs by cell count; no collective is involved. */
int pairhmm_forward_batches_flat(const uint8_t *buf, int64_t buf_bytes,
                                 const int64_t *read_field_off, const int32_t *read_len,
                                 int64_t n_reads, const int64_t *hap_off, const int32_t *hap_len,
                                 int64_t n_haps, const int64_t *batch_read_start,
                                 const int64_t *batch_hap_start, int64_t n_batches,
                                 double *log10_out);

/* A whole pairHMM/test_set-style FILE IMAGE: the batch walk of antidiagsPairHMM.c:375-441 (header line
 * "<num_read> <num_haplotypes>", read lines of five whitespace-separated fields with read length
 * (line length - 4) / 5 (:418), haplotype lines) is reproduced on the GPU (pairhmm_parse.cu), so the
 * caller builds no index arrays and does not need to know the number of results beforehand:
 * *log10_out points at *n_out values, batches in file order, read-major inside a batch (the reference's
 * output order :459-461), and *batch_pairs at *n_batches counts (reads x haplotypes of each batch: what a
 * driver needs to print "#batch: %d" between them).  Both arrays live in pinned host memory OWNED BY THE
 * LIBRARY and stay valid until the next PairHMM call or agx_shutdown().  *incomplete != 0 when the file
 * ends inside a last batch: that batch is dropped and the reference's message is "Error reading
 * haplotypes." (*incomplete = 2: its haplotype cursor runs ahead of its read cursor, :388-396, so this is
 * what it prints whenever the batch has haplotypes) or "Error reading reads." (*incomplete = 1: a batch
 * without haplotypes, :411-414).
 * AGX_ERANGE when a line exceeds the reference's 5000-byte line buffer (:353) or a read length falls
 * outside [1, 8192].  With several GPUs bound and an image of at least 8 MiB per GPU the image is cut at
 * header lines into one range of whole batches per GPU (no collective: every GPU parses and scores its own
 * range; the cuts are verified to be batch boundaries, else the first GPU takes the whole image). */
int pairhmm_forward_file_image(const uint8_t *image, int64_t image_bytes, const double **log10_out,
                               int64_t *n_out, const int32_t **batch_pairs, int64_t *n_batches,
                               int32_t *incomplete);

/* Device-resident variant for one GPU (all pointers device pointers on `device`, work ordered on
 * `stream`).  d_read_batch[r] is the batch of read r; d_read_out_off[r] is the index in
 * d_log10_out of (read r, first haplotype of its batch); n_pairs = sum_b nr_b*nh_b.
 * The call reads a small class histogram back (one stream synchronisation) to size its grids;
 * with fp64_rescue != 0 it synchronises once more to learn how many pairs need the FP64 re-run
 * (with fp64_rescue == 0 such pairs are left as NaN). */
int pairhmm_forward_batches_device(int32_t device, const uint8_t *d_buf, int64_t buf_bytes,
                                   const int64_t *d_read_field_off, const int32_t *d_read_len,
                                   const int32_t *d_read_batch, const int64_t *d_read_out_off,
                                   int64_t n_reads, const int64_t *d_hap_off,
                                   const int32_t *d_hap_len, int64_t n_haps,
                                   const int64_t *d_batch_hap_start, int64_t n_batches,
                                   int64_t n_pairs, int32_t fp64_rescue, double *d_log10_out,
                                   void *stream);

/* The same for several device-resident shards, one per GPU (see sw_score_shards_device); the fields are the
 * arguments of pairhmm_forward_batches_device. */
typedef struct {
    int32_t device;
    const uint8_t *d_buf;
    int64_t buf_bytes;
    const int64_t *d_read_field_off;
    const int32_t *d_read_len;
    const int32_t *d_read_batch;
    const int64_t *d_read_out_off;
    int64_t n_reads;
    const int64_t *d_hap_off;
    const int32_t *d_hap_len;
    int64_t n_haps;
    const int64_t *d_batch_hap_start;
    int64_t n_batches;
    int64_t n_pairs;
    double *d_log10_out;
} agx_hmm_shard;
int pairhmm_forward_shards_device(const agx_hmm_shard *shards, int32_t n_shards, int32_t fp64_rescue);

/* Pairs the FP64 kernel re-ran in the last PairHMM call on `device` that read the count back (the host entry
 * points and the device variants with fp64_rescue != 0); -1 when unknown. */
int64_t agx_pairhmm_rescue_count(int32_t device);

/* PairHMM numeric mode (process-wide).  0 (default) = the reference's semantics (mismatch prior
 * Qr, antidiagsPairHMM.c:111-113).  Bit 0 = corrected GATK semantics (mismatch prior Qr/3); bit 1 (with bit 0)
 * = GATK's base-quality floor as well (base qualities below 6 are read as 6).  Reported separately and never
 * used for parity claims. */
int agx_pairhmm_set_gatk_mode(int32_t on);

/* Force every PairHMM pair through the FP64 kernel (debug / accuracy studies). */
int agx_pairhmm_set_force_fp64(int32_t on);

#ifdef __cplusplus
}
#endif
#endif /* AGX_H */
