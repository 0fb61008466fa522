/*
 * sw_align.c -- CPU definition of the alignment END CELL, START CELL and CIGAR of a Smith-Waterman pair.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * The reference prints the score alone (its README: traceback was left out), so there is no reference OUTPUT to
 * be equal to.  What the reference does fix is WHICH cell its running maximum comes from:
 *   antidiagonalSmithWaterman.c:335   max = val > max ? val : max;      (strict: the first cell wins)
 *   :270-347                          cells are visited anti-diagonal by anti-diagonal (iy + ix ascending), inside
 *                                     one anti-diagonal with ix ascending (iy descending)
 *   :229-244                          ix walks the SHORTER line (sx), line 1 when both are equally long
 * The end cell below is that cell.  It is pinned to the reference's own scan by oracle/_ref/sw_antidiag_ends, the
 * reference source with position bookkeeping added beside line 335 by sed (oracle/Makefile; no arithmetic touched).
 *
 * The traceback is this repository's definition (nothing in the reference to follow).  It uses the D matrix
 * alone: at a cell with D > 0, in this order,
 *   1. diagonal        D[i][j] == D[i-1][j-1] + subst(a_i, b_j)                           -> M
 *   2. for k = 1, 2, ...
 *        a. gap of k symbols of a   D == D[a index - k][same b] + gap_open + k*gap_extend   -> I (k times)
 *        b. gap of k symbols of b   D == D[same a][b index - k] + gap_open + k*gap_extend   -> D (k times)
 * and stops at the first cell with D == 0.  Because Q[i][j] = max_k D[i][j-k] + go + k*ge (and P likewise), one of
 * the three always applies, and the path it spells scores exactly the Smith-Waterman score.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

static inline int32_t max2(int32_t x, int32_t y) { return x > y ? x : y; }

/* D matrix, (lb+1) rows x (la+1) columns, rows follow b; same recurrence as oracle_sw_score (:309-335) */
static int32_t *fill(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb, int32_t match, int32_t mismatch,
                     int32_t gap_open, int32_t gap_extend)
{
    const size_t w = (size_t)la + 1;
    int32_t *H = calloc((size_t)(lb + 1) * w, sizeof(int32_t));
    int32_t *vgap = malloc(w * sizeof(int32_t));
    if (!H || !vgap) { free(H); free(vgap); return NULL; }
    const int32_t NEG = -(1 << 29);
    const int32_t first_gap = gap_open + gap_extend;
    for (int32_t j = 0; j <= la; j++) vgap[j] = NEG;
    for (int32_t i = 1; i <= lb; i++) {
        int32_t hgap = NEG;
        int32_t *row = H + (size_t)i * w;
        const int32_t *up = row - w;
        for (int32_t j = 1; j <= la; j++) {
            const int32_t p = max2(up[j] + first_gap, vgap[j] + gap_extend);
            const int32_t q = max2(row[j - 1] + first_gap, hgap + gap_extend);
            const int32_t s = up[j - 1] + (b[i - 1] == a[j - 1] ? match : mismatch);
            row[j] = max2(max2(p, q), max2(s, 0));
            vgap[j] = p;
            hgap = q;
        }
    }
    free(vgap);
    return H;
}

int32_t oracle_sw_align(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb, int32_t match,
                        int32_t mismatch, int32_t gap_open, int32_t gap_extend, int32_t coords[4],
                        uint32_t *cigar, int32_t cigar_cap, int32_t *n_ops_out)
{
    coords[0] = coords[1] = coords[2] = coords[3] = -1;
    if (n_ops_out) *n_ops_out = 0;
    if (la <= 0 || lb <= 0) return 0;
    int32_t *H = fill(a, la, b, lb, match, mismatch, gap_open, gap_extend);
    if (!H) return INT32_MIN;
    const size_t w = (size_t)la + 1;
    /* the reference's scan: sx (ix) is line 2 only when line 1 is strictly longer (:229) */
    const int a_is_sx = !(la > lb);
    int32_t best = 0, ea = 0, eb = 0;          /* 1-based end cell: ea in a, eb in b */
    int64_t best_key = 0;
    for (int32_t i = 1; i <= lb; i++)
        for (int32_t j = 1; j <= la; j++) {
            const int32_t v = H[(size_t)i * w + j];
            const int64_t key = ((int64_t)(i + j) << 32) | (uint32_t)(a_is_sx ? j : i);   /* visiting order */
            if (v > best || (v == best && v > 0 && key < best_key)) { best = v; ea = j; eb = i; best_key = key; }
        }
    if (best == 0) { free(H); return 0; }
    coords[1] = ea - 1;
    coords[3] = eb - 1;
    /* traceback; ops are collected end -> start, then reversed and run-length merged */
    const int32_t cap = la + lb + 2;
    uint8_t *ops = malloc((size_t)cap);
    int32_t n = 0, i = eb, j = ea;
    while (i > 0 && j > 0) {
        const int32_t h = H[(size_t)i * w + j];
        if (h == 0) break;
        if (h == H[(size_t)(i - 1) * w + j - 1] + (b[i - 1] == a[j - 1] ? match : mismatch)) {
            ops[n++] = 0; --i; --j;
            continue;
        }
        int found = 0;
        for (int32_t k = 1; k <= j || k <= i; k++) {
            if (k <= j && h == H[(size_t)i * w + j - k] + gap_open + k * gap_extend) {
                for (int32_t t = 0; t < k; t++) ops[n++] = 1;
                j -= k; found = 1; break;
            }
            if (k <= i && h == H[(size_t)(i - k) * w + j] + gap_open + k * gap_extend) {
                for (int32_t t = 0; t < k; t++) ops[n++] = 2;
                i -= k; found = 1; break;
            }
        }
        if (!found) { free(ops); free(H); return INT32_MIN; }   /* cannot happen */
    }
    coords[0] = j;       /* first aligned symbol, 0-based */
    coords[2] = i;
    int32_t runs = 0;
    for (int32_t t = n - 1; t >= 0;) {
        int32_t u = t;
        while (u >= 0 && ops[u] == ops[t]) --u;
        if (cigar && runs < cigar_cap) cigar[runs] = ((uint32_t)(t - u) << 4) | ops[t];
        ++runs;
        t = u;
    }
    if (n_ops_out) *n_ops_out = runs;
    free(ops);
    free(H);
    return best;
}

int32_t oracle_sw_ends(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb, int32_t match, int32_t mismatch,
                       int32_t gap_open, int32_t gap_extend, int32_t ends[2])
{
    /* the END CELL alone with two rolling rows (long pairs): same recurrence, same visiting-order rule as
     * oracle_sw_align -- a later cell replaces the kept one only with a larger value, or with the same value on an
     * earlier anti-diagonal / the same anti-diagonal and a smaller ix */
    ends[0] = ends[1] = -1;
    if (la <= 0 || lb <= 0) return 0;
    const size_t w = (size_t)la + 1;
    int32_t *h = calloc(w, sizeof(int32_t)), *vgap = malloc(w * sizeof(int32_t));
    if (!h || !vgap) { free(h); free(vgap); return INT32_MIN; }
    const int32_t NEG = -(1 << 29), first_gap = gap_open + gap_extend;
    const int a_is_sx = !(la > lb);
    for (int32_t j = 0; j <= la; j++) vgap[j] = NEG;
    int32_t best = 0;
    int64_t best_key = 0;
    for (int32_t i = 1; i <= lb; i++) {
        int32_t hgap = NEG, diag = h[0], left = 0;
        for (int32_t j = 1; j <= la; j++) {
            const int32_t up = h[j];
            const int32_t p = max2(up + first_gap, vgap[j] + gap_extend);
            const int32_t q = max2(left + first_gap, hgap + gap_extend);
            const int32_t s = diag + (b[i - 1] == a[j - 1] ? match : mismatch);
            const int32_t v = max2(max2(p, q), max2(s, 0));
            if (v >= best && v > 0) {
                const int64_t key = ((int64_t)(i + j) << 32) | (uint32_t)(a_is_sx ? j : i);
                if (v > best || key < best_key) { best = v; best_key = key; ends[0] = j - 1; ends[1] = i - 1; }
            }
            vgap[j] = p; hgap = q; diag = up; h[j] = v; left = v;
        }
    }
    free(h); free(vgap);
    return best;
}

int32_t oracle_sw_cigar_score(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb, int32_t match,
                              int32_t mismatch, int32_t gap_open, int32_t gap_extend, const int32_t coords[4],
                              const uint32_t *cigar, int32_t n_ops)
{
    /* independent re-scoring of a CIGAR: walks the path and adds up what it spells; INT32_MIN when the path does
     * not run exactly from (coords[0], coords[2]) to (coords[1], coords[3]) */
    int32_t i = coords[0], j = coords[2], s = 0;
    for (int32_t t = 0; t < n_ops; t++) {
        const int32_t len = (int32_t)(cigar[t] >> 4), op = (int32_t)(cigar[t] & 15);
        if (len <= 0) return INT32_MIN;
        if (op == 0) {
            for (int32_t k = 0; k < len; k++, i++, j++) {
                if (i >= la || j >= lb) return INT32_MIN;
                s += a[i] == b[j] ? match : mismatch;
            }
        } else if (op == 1) { i += len; s += gap_open + len * gap_extend; }
        else if (op == 2) { j += len; s += gap_open + len * gap_extend; }
        else return INT32_MIN;
    }
    if (i - 1 != coords[1] || j - 1 != coords[3]) return INT32_MIN;
    return s;
}
