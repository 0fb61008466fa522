/*
 * oracle.h -- CPU restatement of the reference's two dynamic-programming hot paths.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load liboracle.so.  The product (libagx.so, drivers/) never links,
 * loads or calls anything in this directory and has no CPU fallback.
 *
 * Parity pins (see DESIGN.md "Oracle"):
 *   - pairHMM/test_set/test.out (the reference's only committed golden): -4.485565
 *   - outputs of the UNMODIFIED reference C programs compiled by `make -C oracle ref`
 *     (oracle/_ref/), recorded under tests/golden/ by tests/golden/make_golden.py.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 */
#ifndef AGX_ORACLE_H
#define AGX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Score-only affine-gap local alignment over RAW BYTES (byte equality, '\n' is a symbol).
 * Restates smithWaterman/antidiagonalSmithWaterman.c:290-335 (boundary init + P/Q/D recurrence)
 * as a row-major fill; the reference's rolling 3-anti-diagonal store (:96-184) is pure storage.
 * gap_open/gap_extend follow the reference's sign convention (negative numbers; the first gap
 * base costs gap_open + gap_extend, :313, :321).  corner_out (may be NULL) receives D[la][lb]. */
int32_t oracle_sw_score(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb,
                        int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                        int32_t *corner_out);

/* The same recurrence evaluated tile by tile on `threads` host threads (sw_blocked.c; pthreads, threads <= 0:
 * one per online core): what makes a 1 Mbp x 1 Mbp expected score (BASELINE configs[4]) computable.
 * Pinned to oracle_sw_score() by tests/test_oracle.py.  tile <= 0 picks 4096. */
int32_t oracle_sw_score_blocked(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb,
                                int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                                int32_t tile, int32_t threads);

/* Alignment END CELL, START CELL and CIGAR of one pair (sw_align.c; a = line 1, b = line 2, RAW BYTES).
 * Returns the score (== oracle_sw_score).  coords = {a_start, a_end, b_start, b_end}: 0-based indices of the first
 * and last aligned symbol of each sequence, all -1 when the score is 0.  The END is the cell the reference's
 * running maximum comes from (antidiagonalSmithWaterman.c:335, strict `>`, cells visited by anti-diagonals
 * :270-347 with ix -- the shorter line, line 1 on ties, :229-244 -- ascending); pinned to the reference's own scan
 * by oracle/_ref/sw_antidiag_ends.  START and CIGAR follow the traceback rule stated in sw_align.c (this
 * repository's definition; the reference has none).  cigar[k] = length << 4 | op with op 0 = M (one symbol of
 * each), 1 = I (symbols of a only), 2 = D (symbols of b only), start -> end; *n_ops_out = number of runs (may exceed
 * cigar_cap, then only cigar_cap are written). */
int32_t oracle_sw_align(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb, int32_t match,
                        int32_t mismatch, int32_t gap_open, int32_t gap_extend, int32_t coords[4],
                        uint32_t *cigar, int32_t cigar_cap, int32_t *n_ops_out);

/* The END CELL alone (ends = {a_end, b_end}, 0-based, -1 -1 when the score is 0) with two rolling rows: what checks
 * long pairs, whose full matrix oracle_sw_align could not hold.  Returns the score. */
int32_t oracle_sw_ends(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb, int32_t match, int32_t mismatch,
                       int32_t gap_open, int32_t gap_extend, int32_t ends[2]);

/* Independent re-scoring of a CIGAR (walks the path, adds up substitutions and gaps): INT32_MIN when the path
 * does not run exactly from the start to the end coordinates. */
int32_t oracle_sw_cigar_score(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb, int32_t match,
                              int32_t mismatch, int32_t gap_open, int32_t gap_extend, const int32_t coords[4],
                              const uint32_t *cigar, int32_t n_ops);

/* File-level restatement of antidiagonalSmithWaterman.c:205-227, 348 (header = number of LINES to
 * consume, fgets into a `line_buf`-byte buffer so longer lines split, trailing '\n' kept as a
 * symbol, EOF mid-pair stops).  line_buf = 1000 reproduces the unmodified program.
 * Returns the number of scores written (<= cap), or -1 if the file cannot be opened / is empty.
 * header_out (may be NULL) receives atoi(first line). */
int64_t oracle_sw_file(const char *path, int32_t line_buf, int32_t *scores, int64_t cap,
                       int32_t *header_out);

/* Phred+33 character -> error probability, pairHMM/antidiagsPairHMM.c:104-107
 * (= pairHMMmatrix.c:25-28): pow(10.0, -(c - 33.0) * 0.1) in double with the host libm. */
double oracle_pairhmm_prob(uint8_t phred33);

/* PairHMM forward in double, one (read, haplotype) pair.
 * Restates pairHMMmatrix.c:41-56 (init + M/X/Y recurrences, identical maths to
 * antidiagsPairHMM.c:157-202), p() :32-34 (mismatch prior is Qr, NOT Qr/3 -- the reference's
 * quirk HMM-Q1, reproduced), mm() :36-38, and the final sum + log10 of :59-66
 * (= antidiagsPairHMM.c:206-212, 242).  gatk_mode bit 0 switches the mismatch prior to Qr/3,
 * bit 1 adds GATK's base-quality floor of 6 (the corrected semantics the north star reports separately; NOT
 * the reference). */
double oracle_pairhmm_forward(const uint8_t *bases, const uint8_t *q, const uint8_t *qi,
                              const uint8_t *qd, const uint8_t *qg, int32_t read_len,
                              const uint8_t *hap, int32_t hap_len, int32_t gatk_mode);

/* File-level restatement of antidiagsPairHMM.c:371-491 / pairHMMmatrix.c:167-318: batches of
 * "<num_read> <num_haplotypes>", read lines of five equal-length fields, haplotype lines;
 * read length inferred as (strlen(line)-4)/5 (:418).  Results read-major, haplotype-minor.
 * Returns the number of results written (<= cap) or -1 on open failure. n_batches_out may be NULL. */
int64_t oracle_pairhmm_file(const char *path, double *out, int64_t cap, int32_t *n_batches_out);

#ifdef __cplusplus
}
#endif
#endif
