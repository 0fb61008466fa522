/*
 * oracle.c -- CPU restatement of the reference hot paths.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Written from the maths in SURVEY.md section 8(a), not from the reference's text: plain row-major
 * fills with two rolling rows instead of the reference's 3-anti-diagonal store.  Validated against
 * the compiled reference programs (oracle/_ref, `make ref`) by tests/test_oracle.py and pinned by
 * tests/golden/.
 */
#define _POSIX_C_SOURCE 200809L
#include "oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* "minus infinity" that survives additions, as antidiagonalSmithWaterman.c:86-88 does for INT_MIN */
#define NEG_INF INT_MIN

static inline int32_t add_inf(int32_t v, int32_t d) { return v == NEG_INF ? NEG_INF : v + d; }
static inline int32_t max2(int32_t x, int32_t y) { return x > y ? x : y; }

int32_t oracle_sw_score(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb,
                        int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                        int32_t *corner_out)
{
    /* rows follow b (the reference's sy), columns follow a (sx); the score is symmetric. */
    int32_t *vgap = malloc((size_t)(la + 1) * sizeof(int32_t)); /* P: gap running down a column */
    int32_t *best = malloc((size_t)(la + 1) * sizeof(int32_t)); /* D of the previous row        */
    int32_t top = 0, corner = 0;
    if (!vgap || !best) { free(vgap); free(best); return INT_MIN; }

    /* row 0 (:290-297): P = -inf, D = 0 */
    for (int32_t j = 0; j <= la; j++) { vgap[j] = NEG_INF; best[j] = 0; }

    const int32_t first_gap = gap_open + gap_extend; /* :313, :321 */
    for (int32_t i = 1; i <= lb; i++) {
        /* column 0 (:299-306): Q = -inf, D = 0 */
        int32_t hgap = NEG_INF;
        int32_t left = 0;          /* D[i][j-1]   */
        int32_t diag = best[0];    /* D[i-1][j-1] */
        best[0] = 0;
        for (int32_t j = 1; j <= la; j++) {
            int32_t up = best[j];                                             /* D[i-1][j] */
            int32_t p = max2(add_inf(up, first_gap), add_inf(vgap[j], gap_extend));   /* :309-315 */
            int32_t q = max2(add_inf(left, first_gap), add_inf(hgap, gap_extend));    /* :317-323 */
            int32_t s = diag + (b[i - 1] == a[j - 1] ? match : mismatch);             /* :332 */
            int32_t d = max2(max2(p, q), max2(s, 0));                                 /* :333 */
            if (d > top) top = d;                                                     /* :335 */
            vgap[j] = p;
            hgap = q;
            diag = up;
            best[j] = d;
            left = d;
        }
    }
    corner = (la > 0 && lb > 0) ? best[la] : 0;
    if (corner_out) *corner_out = corner;
    free(vgap);
    free(best);
    return top;
}

int64_t oracle_sw_file(const char *path, int32_t line_buf, int32_t *scores, int64_t cap,
                       int32_t *header_out)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char *l1 = malloc((size_t)line_buf), *l2 = malloc((size_t)line_buf);
    int64_t n = 0;
    if (!l1 || !l2 || !fgets(l1, line_buf, f)) { free(l1); free(l2); fclose(f); return -1; }
    int header = atoi(l1);                                   /* :209 */
    if (header_out) *header_out = header;
    for (int i = 0; i < header; i += 2) {                    /* :216 -- header counts LINES */
        if (!fgets(l1, line_buf, f)) break;                  /* :219 */
        if (!fgets(l2, line_buf, f)) break;                  /* :223 (reference echoes line 1) */
        int32_t n1 = (int32_t)strlen(l1), n2 = (int32_t)strlen(l2); /* '\n' included, :229-244 */
        if (n < cap)
            scores[n] = oracle_sw_score((const uint8_t *)l1, n1, (const uint8_t *)l2, n2,
                                        1, -1, -3, -1, NULL); /* :40-43 */
        n++;
    }
    free(l1); free(l2);
    fclose(f);
    return n;
}

double oracle_pairhmm_prob(uint8_t c)
{
    return pow(10.0, -((char)c - 33.0) * 0.1);
}

double oracle_pairhmm_forward(const uint8_t *bases, const uint8_t *q, const uint8_t *qi,
                              const uint8_t *qd, const uint8_t *qg, int32_t read_len,
                              const uint8_t *hap, int32_t hap_len, int32_t gatk_mode)
{
    const size_t w = (size_t)hap_len + 1;
    double *buf = calloc(6 * w, sizeof(double));
    if (!buf) return NAN;
    double *m0 = buf, *x0 = buf + w, *y0 = buf + 2 * w;     /* row i-1 */
    double *m1 = buf + 3 * w, *x1 = buf + 4 * w, *y1 = buf + 5 * w; /* row i */

    /* row 0: M = X = 0, Y = (DBL_MAX/16)/hap_len for j = 0..hap_len (pairHMMmatrix.c:43-46) */
    const double init = DBL_MAX / 16 / (double)hap_len;
    for (int32_t j = 0; j <= hap_len; j++) y0[j] = init;

    for (int32_t i = 1; i <= read_len; i++) {
        /* gatk_mode bit 1: GATK reads base qualities below 6 as 6 (NOT the reference) */
        const uint8_t qc = ((gatk_mode & 2) && (char)q[i - 1] < 33 + 6) ? 33 + 6 : q[i - 1];
        const double Qr = oracle_pairhmm_prob(qc);
        const double Qi = oracle_pairhmm_prob(qi[i - 1]);
        const double Qd = oracle_pairhmm_prob(qd[i - 1]);
        const double Qg = oracle_pairhmm_prob(qg[i - 1]);
        const char r = (char)bases[i - 1];
        m1[0] = x1[0] = y1[0] = 0.0;                         /* column 0 stays zero (memset) */
        for (int32_t j = 1; j <= hap_len; j++) {
            const char h = (char)hap[j - 1];
            /* p(): :32-34 -- match prior 1-Qr; mismatch prior Qr (no /3: quirk HMM-Q1) */
            double prior = (r == h || r == 'N' || h == 'N') ? 1 - Qr : ((gatk_mode & 1) ? Qr / 3 : Qr);
            /* :51-53, same association order as the reference expression */
            m1[j] = prior * ((1 - (Qi + Qd)) * m0[j - 1] + (1 - Qg) * (x0[j - 1] + y0[j - 1]));
            x1[j] = m0[j] * Qi + x0[j] * Qg;
            y1[j] = m1[j - 1] * Qd + y1[j - 1] * Qg;
        }
        double *t;
        t = m0; m0 = m1; m1 = t;
        t = x0; x0 = x1; x1 = t;
        t = y0; y0 = y1; y1 = t;
    }
    /* final sum over the last row in increasing j (:59-66); row read_len now lives in m0/x0 */
    double l = 0;
    for (int32_t j = 1; j <= hap_len; j++) l += m0[j] + x0[j];
    free(buf);
    return log10(l) - log10(DBL_MAX / 16);
}

#define HMM_LINE 5001 /* MAX_READ_LEN*5+1, antidiagsPairHMM.c:8, :353 */

int64_t oracle_pairhmm_file(const char *path, double *out, int64_t cap, int32_t *n_batches_out)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char line[HMM_LINE];
    int64_t n = 0;
    int32_t batches = 0;
    int nr = 0, nh = 0;                                      /* :345-346: declared once, so a header that sscanf
                                                                cannot (fully) parse keeps the previous count(s) */
    while (fgets(line, sizeof line, f)) {                    /* :375 */
        sscanf(line, "%d %d", &nr, &nh);                     /* :378 */
        char **reads = calloc((size_t)(nr > 0 ? nr : 1), sizeof(char *));
        char **haps = calloc((size_t)(nh > 0 ? nh : 1), sizeof(char *));
        int ok = reads && haps;
        for (int i = 0; ok && i < nr; i++) {
            if (!fgets(line, sizeof line, f)) { ok = 0; break; }
            line[strcspn(line, "\n")] = '\0';                /* :417 */
            reads[i] = strdup(line);
        }
        for (int i = 0; ok && i < nh; i++) {
            if (!fgets(line, sizeof line, f)) { ok = 0; break; }
            line[strcspn(line, "\n")] = '\0';                /* :399 */
            haps[i] = strdup(line);
        }
        for (int i = 0; ok && i < nr; i++) {
            int len = ((int)strlen(reads[i]) - 4) / 5;       /* :418 */
            if (len < 0) len = 0;
            size_t cap_f = strlen(reads[i]) + 1;
            char *fld = malloc(5 * cap_f);
            char *b = fld, *m = fld + cap_f, *in = fld + 2 * cap_f, *d = fld + 3 * cap_f,
                 *g = fld + 4 * cap_f;
            b[0] = m[0] = in[0] = d[0] = g[0] = '\0';
            sscanf(reads[i], "%s %s %s %s %s", b, m, in, d, g); /* :101 */
            for (int j = 0; j < nh; j++) {
                double v = oracle_pairhmm_forward((uint8_t *)b, (uint8_t *)m, (uint8_t *)in,
                                                  (uint8_t *)d, (uint8_t *)g, len,
                                                  (uint8_t *)haps[j], (int32_t)strlen(haps[j]), 0);
                if (n < cap) out[n] = v;
                n++;
            }
            free(fld);
        }
        for (int i = 0; i < nr && reads; i++) free(reads[i]);
        for (int i = 0; i < nh && haps; i++) free(haps[i]);
        free(reads); free(haps);
        if (!ok) break;
        batches++;
    }
    if (n_batches_out) *n_batches_out = batches;
    fclose(f);
    return n;
}
