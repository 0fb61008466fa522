/*
 * sw_blocked.c -- the SAME recurrence as oracle_sw_score() (oracle.c; antidiagonalSmithWaterman.c:290-335),
 * evaluated tile by tile on several host threads so that a 1 Mbp x 1 Mbp matrix (10^12 cells,
 * BASELINE configs[4]) gets an independently computed score in minutes instead of core-hours.
 * TEST INFRASTRUCTURE ONLY (see oracle.h): it produces the committed expected scores under
 * tests/golden/ and checks the long-alignment GPU kernels; nothing in the product loads it.
 *
 * The matrix is cut into tile x tile blocks.  Blocks on one anti-diagonal of the block grid are
 * independent (a pool of pthreads takes them from a shared counter); a block reads
 *     the bottom row (D, P) of the block above      -> row[]  / vrow[]   (indexed by column, in place)
 *     the right column (D, Q) of the block to its left -> col[] / hcol[] (indexed by row, in place)
 *     D of the cell diagonally above-left of its first cell -> corner[][]
 * and leaves its own bottom row / right column / bottom-right D there for its neighbours.
 * "-inf" is INT_MIN / 2: additions of gap penalties cannot wrap for any matrix this file is used on
 * (|penalty| * (la + lb) < 2^30), and a value that low never wins a max against D >= 0 -- the same
 * argument as antidiagonalSmithWaterman.c:86-88 (sum_with_infinity) without the branch.
 * Pinned to oracle_sw_score() on random pairs and tile sizes by tests/test_oracle.py.
 */
#define _POSIX_C_SOURCE 200809L
#include "oracle.h"

#include <limits.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define NEG (INT_MIN / 2)

static inline int32_t mx(int32_t x, int32_t y) { return x > y ? x : y; }

struct blocked {
    const uint8_t *a, *b;
    int64_t la, lb, nbi, nbj, tile;
    int32_t match, mismatch, first_gap, gap_extend;
    int32_t *row, *vrow, *col, *hcol, *corner;
    atomic_llong next;            /* next block of the current anti-diagonal */
    pthread_barrier_t bar;
};

/* one block: rows r0+1 .. r1, columns c0+1 .. c1; returns the largest D inside it */
static int32_t block(struct blocked *g, int64_t bi, int64_t bj)
{
    const int64_t r0 = bi * g->tile, r1 = (r0 + g->tile < g->lb) ? r0 + g->tile : g->lb;
    const int64_t c0 = bj * g->tile, c1 = (c0 + g->tile < g->la) ? c0 + g->tile : g->la;
    const int32_t first_gap = g->first_gap, gap_extend = g->gap_extend, match = g->match, mismatch = g->mismatch;
    int32_t *row = g->row, *vrow = g->vrow;
    const uint8_t *a = g->a;
    int32_t above_left = g->corner[bi * (g->nbj + 1) + bj];  /* D[r0][c0] */
    int32_t best = 0;
    for (int64_t i = r0 + 1; i <= r1; i++) {
        int32_t left = g->col[i], hgap = g->hcol[i];          /* D[i][c0], Q[i][c0] */
        int32_t diag = above_left;                            /* D[i-1][c0] */
        above_left = left;
        const uint8_t bc = g->b[i - 1];
        for (int64_t j = c0 + 1; j <= c1; j++) {
            const int32_t up = row[j];                                    /* D[i-1][j]   */
            const int32_t p = mx(up + first_gap, vrow[j] + gap_extend);   /* :309-315    */
            const int32_t q = mx(left + first_gap, hgap + gap_extend);    /* :317-323    */
            const int32_t s = diag + (bc == a[j - 1] ? match : mismatch); /* :332        */
            const int32_t dd = mx(mx(p, q), mx(s, 0));                    /* :333        */
            best = mx(best, dd);                                          /* :335        */
            vrow[j] = p;
            hgap = q;
            diag = up;
            row[j] = dd;
            left = dd;
        }
        g->col[i] = left;
        g->hcol[i] = hgap;
    }
    g->corner[(bi + 1) * (g->nbj + 1) + (bj + 1)] = row[c1];  /* D[r1][c1] */
    return best;
}

struct worker { struct blocked *g; int32_t best; int id; };

static void *work(void *arg)
{
    struct worker *w = arg;
    struct blocked *g = w->g;
    int32_t best = 0;
    for (int64_t d = 0; d < g->nbi + g->nbj - 1; d++) {
        const int64_t bi_lo = d < g->nbj ? 0 : d - g->nbj + 1;
        const int64_t bi_hi = d < g->nbi ? d : g->nbi - 1;
        for (;;) {
            const int64_t k = atomic_fetch_add(&g->next, 1);
            if (bi_lo + k > bi_hi) break;
            best = mx(best, block(g, bi_lo + k, d - (bi_lo + k)));
        }
        /* everyone is done with anti-diagonal d before the counter is reset and d + 1 starts */
        pthread_barrier_wait(&g->bar);
        if (w->id == 0) atomic_store(&g->next, 0);
        pthread_barrier_wait(&g->bar);
    }
    w->best = best;
    return NULL;
}

int32_t oracle_sw_score_blocked(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb,
                                int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend,
                                int32_t tile, int32_t threads)
{
    if (la <= 0 || lb <= 0) return 0;
    struct blocked g;
    memset(&g, 0, sizeof g);
    g.a = a; g.b = b; g.la = la; g.lb = lb;
    g.tile = tile < 1 ? 4096 : tile;
    g.nbj = (la + g.tile - 1) / g.tile;
    g.nbi = (lb + g.tile - 1) / g.tile;
    g.match = match; g.mismatch = mismatch;
    g.first_gap = gap_open + gap_extend;                         /* :313, :321 */
    g.gap_extend = gap_extend;
    g.row = malloc((size_t)(la + 1) * sizeof(int32_t));          /* D[i_done][j] */
    g.vrow = malloc((size_t)(la + 1) * sizeof(int32_t));         /* P[i_done][j] */
    g.col = malloc((size_t)(lb + 1) * sizeof(int32_t));          /* D[i][j_done] */
    g.hcol = malloc((size_t)(lb + 1) * sizeof(int32_t));         /* Q[i][j_done] */
    g.corner = calloc((size_t)(g.nbi + 1) * (size_t)(g.nbj + 1), sizeof(int32_t));
    int32_t top = INT_MIN;
    if (g.row && g.vrow && g.col && g.hcol && g.corner) {
        /* row 0 (:290-297): P = -inf, D = 0;  column 0 (:299-306): Q = -inf, D = 0 */
        for (int64_t j = 0; j <= la; j++) { g.row[j] = 0; g.vrow[j] = NEG; }
        for (int64_t i = 0; i <= lb; i++) { g.col[i] = 0; g.hcol[i] = NEG; }
        int n = threads > 0 ? threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
        const int64_t widest = g.nbi < g.nbj ? g.nbi : g.nbj;
        if (n > widest) n = (int)widest;
        if (n < 1) n = 1;
        if (n > 256) n = 256;
        atomic_init(&g.next, 0);
        pthread_barrier_init(&g.bar, NULL, (unsigned)n);
        pthread_t th[256];
        struct worker w[256];
        for (int k = 0; k < n; k++) { w[k].g = &g; w[k].best = 0; w[k].id = k; }
        for (int k = 1; k < n; k++) pthread_create(&th[k], NULL, work, &w[k]);
        work(&w[0]);
        top = w[0].best;
        for (int k = 1; k < n; k++) { pthread_join(th[k], NULL); top = mx(top, w[k].best); }
        pthread_barrier_destroy(&g.bar);
    }
    free(g.row); free(g.vrow); free(g.col); free(g.hcol); free(g.corner);
    return top;
}
