/*
 * smithWatermanAlign -- the reference's Smith-Waterman command line with the alignment itself on every line.
 *
 *   usage:  smithWatermanAlign <file_path>        (same argument as antidiagonalSmithWaterman.c:190-193)
 *
 * The reference prints the score alone (its README: traceback was left out).  This program keeps its input
 * format and line structure -- "line_num: %d" (:209-210), one "Score: %d" per pair in file order (:348), the
 * dangling first line of an incomplete last pair echoed (:223-227), "elapsed %f" (:351-352) -- and appends to
 * each score line where the alignment lies and how it reads:
 *
 *     Score: 23 a[4,30] b[0,27] 12M1I14M
 *
 * a = line 1 of the pair, b = line 2, 0-based inclusive indices of the first and last aligned symbol ("-" when
 * the score is 0); the END is the cell the reference's running maximum comes from (:335, its visiting order),
 * start and CIGAR (M = one symbol of each line, I = symbols of a only, D = symbols of b only) follow the rule in
 * include/agx.h (sw_align_batch_flat).  Sequences are the reference's fgets() chunks (MAX_LINE_LENGTH 1000, :44;
 * AGX_SW_LINE_BUF=<n> raises it), trailing '\n' kept as a symbol.  The DP, the traceback matrices and the walk
 * all live on the GPU (libagx.so); there is no CPU fallback.  AGX_NUM_GPUS=<n> binds n GPUs (default 1, 0 = all).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "agx.h"

static double seconds(void)
{
    struct timeval tp;
    gettimeofday(&tp, NULL);
    return (double)tp.tv_sec + (double)tp.tv_usec * 1.e-6;
}

int main(int argc, char *argv[])
{
    if (argc != 2) {
        fprintf(stderr, "Usage: %s <file_path>\n", argv[0]);
        return 1;
    }
    FILE *file = fopen(argv[1], "rb");
    if (file == NULL) {
        perror("Error opening file");
        exit(EXIT_FAILURE);
    }
    size_t cap = 1 << 20, size = 0;
    unsigned char *img = malloc(cap);
    for (;;) {
        if (size == cap) { cap *= 2; img = realloc(img, cap); }
        if (!img) { fprintf(stderr, "out of memory\n"); return 1; }
        size_t got = fread(img + size, 1, cap - size, file);
        if (got == 0) break;
        size += got;
    }
    fclose(file);
    if (size == 0) {
        printf("file is empty");
        return 1;
    }
    long line_buf = 1000;                         /* MAX_LINE_LENGTH, antidiagonalSmithWaterman.c:44 */
    const char *env = getenv("AGX_SW_LINE_BUF");
    if (env && atol(env) >= 2) line_buf = atol(env);

    /* fgets() chunks of the image: at most line_buf - 1 bytes, a '\n' ends a chunk and stays in it (:201-244) */
    size_t n_chunks = 0, chunk_cap = 1024;
    int64_t *off = malloc(chunk_cap * sizeof *off);
    int32_t *len = malloc(chunk_cap * sizeof *len);
    for (size_t at = 0; at < size;) {
        size_t lim = size - at < (size_t)line_buf - 1 ? size - at : (size_t)line_buf - 1;
        unsigned char *nl = memchr(img + at, '\n', lim);
        size_t l = nl ? (size_t)(nl - (img + at)) + 1 : lim;
        if (n_chunks == chunk_cap) {
            chunk_cap *= 2;
            off = realloc(off, chunk_cap * sizeof *off);
            len = realloc(len, chunk_cap * sizeof *len);
        }
        if (!off || !len) { fprintf(stderr, "out of memory\n"); return 1; }
        off[n_chunks] = (int64_t)at;
        len[n_chunks] = (int32_t)l;
        n_chunks++;
        at += l;
    }
    char tmp[32];
    size_t c = (size_t)len[0] < sizeof tmp - 1 ? (size_t)len[0] : sizeof tmp - 1;
    memcpy(tmp, img, c);
    tmp[c] = 0;
    const int header = atoi(tmp);                                  /* :209: the number of LINES to consume */
    printf("line_num: %d\n", header);
    double iStart = seconds();

    int64_t want = header > 0 ? ((int64_t)header + 1) / 2 : 0;     /* for (i = 0; i < line_num; i += 2), :216 */
    int64_t have = (int64_t)(n_chunks - 1) / 2;
    int64_t n_pairs = want < have ? want : have;
    int dangling = want > have && (n_chunks - 1) % 2 == 1;         /* EOF after the first line of a pair, :223-227 */

    int n_gpus = 1;
    env = getenv("AGX_NUM_GPUS");
    if (env) n_gpus = atoi(env);
    int32_t *scores = malloc((size_t)(n_pairs + 1) * sizeof *scores);
    int32_t *coords = malloc((size_t)(n_pairs + 1) * 4 * sizeof *coords);
    int64_t *cig_off = malloc((size_t)(n_pairs + 2) * sizeof *cig_off);
    int64_t cig_cap = 8 * n_pairs + 64, cig_total = 0;
    uint32_t *cigar = malloc((size_t)cig_cap * sizeof *cigar);
    if (!scores || !coords || !cig_off || !cigar) { fprintf(stderr, "out of memory\n"); return 1; }
    int rc = agx_init(n_gpus);
    for (int attempt = 0; rc == AGX_OK && attempt < 2; attempt++) {
        /* the reference's scoring constants, :40-43 */
        rc = sw_align_batch_flat(img, (int64_t)size, off + 1, len + 1, n_pairs, 1, -1, -3, -1, scores, coords, cig_off,
                                 cigar, cig_cap, &cig_total);
        if (rc == AGX_ERANGE && cig_total > cig_cap) {             /* more runs than guessed: ask again with room */
            cig_cap = cig_total;
            free(cigar);
            cigar = malloc((size_t)cig_cap * sizeof *cigar);
            if (!cigar) { fprintf(stderr, "out of memory\n"); return 1; }
            rc = AGX_OK;
            continue;
        }
        break;
    }
    if (rc != AGX_OK) {
        fprintf(stderr, "Error: code: %d, reason: %s\n", rc, agx_last_error());
        exit(1);
    }
    for (int64_t p = 0; p < n_pairs; p++) {
        if (scores[p] == 0) {
            printf("Score: 0 - - -\n");
            continue;
        }
        printf("Score: %d a[%d,%d] b[%d,%d] ", scores[p], coords[4 * p], coords[4 * p + 1], coords[4 * p + 2], coords[4 * p + 3]);
        for (int64_t k = cig_off[p]; k < cig_off[p + 1]; k++) printf("%u%c", cigar[k] >> 4, "MID"[cigar[k] & 15]);
        putchar('\n');
    }
    if (dangling) fwrite(img + off[n_chunks - 1], 1, (size_t)len[n_chunks - 1], stdout);
    double iElaps = seconds() - iStart;
    printf("elapsed %f\n", iElaps);
    agx_shutdown();
    free(scores); free(coords); free(cig_off); free(cigar); free(off); free(len); free(img);
    return 0;
}
