/*
 * pairHMM -- drop-in for the reference's PairHMM programs, backed by libagx.so.
 *
 *   usage:  pairHMM <input_file_r> <output_file>      (antidiagsPairHMM.c:314-317)
 *
 * Keeps the reference's command line, input format, output format and exit codes:
 *   - batches "<num_read> <num_haplotypes>", num_read read lines of five space-separated fields
 *     (bases, base / insertion / deletion / gap-continuation qualities, Phred+33), num_haplotypes
 *     haplotype lines (antidiagsPairHMM.c:375-441); read length = (strlen(line) - 4) / 5 (:418);
 *     line buffer 5001 bytes (:353);
 *   - stdout: "#batch: %d" before every batch and once more at EOF (:372), each likelihood "%f"
 *     (:459); output file: one "%f" per (read, haplotype), read-major (:461).
 * The file image is handed to pairhmm_forward_file_image(): the batch walk / field splitting
 * (:375-441), the prior setup, the M/X/Y recurrence and the final log10 sum (:99-117, :120-242) all run
 * on the GPU.  Only a file with a line longer than the reference's 5000-byte fgets() buffer -- which the
 * reference would read in pieces -- is split on the host below and goes through
 * pairhmm_forward_batches_flat() instead.  AGX_NUM_GPUS=<n> limits the devices (default: all);
 * AGX_PAIRHMM_GATK=1 selects the corrected GATK mismatch prior (Qr/3), which is NOT the
 * reference's semantics; AGX_PAIRHMM_FP64=1 sends every pair through the FP64 kernel that keeps
 * the reference's own operation order.  There is no CPU fallback.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "agx.h"

#define LINE_BUF 5001 /* MAX_READ_LEN*5+1 */

typedef struct {
    int64_t *v;
    size_t n, cap;
} vec64;
typedef struct {
    int32_t *v;
    size_t n, cap;
} vec32;

static void push64(vec64 *a, int64_t x)
{
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 1024; a->v = realloc(a->v, a->cap * sizeof *a->v); }
    if (!a->v) { fprintf(stderr, "out of memory\n"); exit(EXIT_FAILURE); }
    a->v[a->n++] = x;
}
static void push32(vec32 *a, int32_t x)
{
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 1024; a->v = realloc(a->v, a->cap * sizeof *a->v); }
    if (!a->v) { fprintf(stderr, "out of memory\n"); exit(EXIT_FAILURE); }
    a->v[a->n++] = x;
}

/* next fgets()-style chunk of the image; returns 0 at EOF.  *l excludes the trailing '\n'. */
static int next_line(const unsigned char *img, size_t size, size_t *pos, size_t *start, size_t *l)
{
    if (*pos >= size) return 0;
    size_t lim = size - *pos < LINE_BUF - 1 ? size - *pos : LINE_BUF - 1;
    const unsigned char *nl = memchr(img + *pos, '\n', lim);
    size_t raw = nl ? (size_t)(nl - (img + *pos)) + 1 : lim;
    *start = *pos;
    *l = nl ? raw - 1 : raw;
    *pos += raw;
    return 1;
}

/* Host-side fgets() walk (antidiagsPairHMM.c:375-441) for files whose lines exceed the line buffer. */
static int split_on_host(const unsigned char *img, size_t size, double **lh_out, int64_t *n_out_p,
                         int32_t **bp_out, int64_t *n_batches_p, const char **err_p)
{
    vec64 rfo = {0}, ho = {0}, brs = {0}, bhs = {0};
    vec32 rl = {0}, hl = {0};
    push64(&brs, 0);
    push64(&bhs, 0);
    size_t pos = 0, start, l;
    const char *err = NULL;
    /* declared once, outside the batch loop, as in the reference (antidiagsPairHMM.c:345-346): a header that
     * sscanf() cannot (fully) parse keeps the previous batch's count(s) */
    int num_read = 0, num_haplotypes = 0;
    while (next_line(img, size, &pos, &start, &l)) {
        char head[64];
        size_t c = l < sizeof head - 1 ? l : sizeof head - 1;
        memcpy(head, img + start, c);
        head[c] = 0;
        sscanf(head, "%d %d", &num_read, &num_haplotypes);
        size_t r_mark = rl.n, h_mark = hl.n;
        for (int i = 0; i < num_read && !err; i++) {
            /* the reference's haplotype cursor runs ahead (:388-396): it is the one that meets EOF first */
            if (!next_line(img, size, &pos, &start, &l)) {
                err = num_haplotypes > 0 ? "Error reading haplotypes.\n" : "Error reading reads.\n";
                break;
            }
            int len_read = ((int)l - 4) / 5;
            if (len_read < 0) len_read = 0;
            /* the five whitespace-separated fields, as sscanf("%s %s %s %s %s") finds them */
            size_t p = start, end = start + l;
            for (int f = 0; f < 5; f++) {
                while (p < end && (img[p] == ' ' || img[p] == '\t')) p++;
                push64(&rfo, (int64_t)p);
                while (p < end && img[p] != ' ' && img[p] != '\t') p++;
            }
            push32(&rl, len_read);
        }
        for (int i = 0; i < num_haplotypes && !err; i++) {
            if (!next_line(img, size, &pos, &start, &l)) { err = "Error reading haplotypes.\n"; break; }
            push64(&ho, (int64_t)start);
            push32(&hl, (int32_t)l);
        }
        if (err) {  /* incomplete batch: drop it, report like the reference, keep earlier batches */
            rl.n = r_mark; rfo.n = 5 * r_mark; hl.n = h_mark; ho.n = h_mark;
            break;
        }
        push64(&brs, (int64_t)rl.n);
        push64(&bhs, (int64_t)hl.n);
    }
    const int64_t n_batches = (int64_t)brs.n - 1;
    int64_t n_out = 0;
    int32_t *bp = malloc((size_t)(n_batches > 0 ? n_batches : 1) * sizeof *bp);
    for (int64_t b = 0; b < n_batches; b++) {
        bp[b] = (int32_t)((brs.v[b + 1] - brs.v[b]) * (bhs.v[b + 1] - bhs.v[b]));
        n_out += bp[b];
    }
    double *lh = malloc((size_t)(n_out > 0 ? n_out : 1) * sizeof *lh);
    int rc = AGX_OK;
    if (n_out > 0)
        rc = pairhmm_forward_batches_flat(img, (int64_t)size, rfo.v, rl.v, (int64_t)rl.n, ho.v, hl.v,
                                          (int64_t)hl.n, brs.v, bhs.v, n_batches, lh);
    free(rfo.v); free(ho.v); free(brs.v); free(bhs.v); free(rl.v); free(hl.v);
    *lh_out = lh;
    *n_out_p = n_out;
    *bp_out = bp;
    *n_batches_p = n_batches;
    *err_p = err;
    return rc;
}

int main(int argc, const char *argv[])
{
    if (argc != 3) {
        fprintf(stderr, "Usage: %s <input_file_r> <output_file>\n", argv[0]);
        return EXIT_FAILURE;
    }
    FILE *in = fopen(argv[1], "rb");
    if (in == NULL) {
        perror("Error opening input file_r");
        return EXIT_FAILURE;
    }
    FILE *out = fopen(argv[2], "w");
    if (out == NULL) {
        perror("Error opening output file");
        fclose(in);
        return EXIT_FAILURE;
    }
    size_t cap = 1 << 20, size = 0;
    unsigned char *img = malloc(cap);
    for (;;) {
        if (size == cap) { cap *= 2; img = realloc(img, cap); }
        if (!img) { fprintf(stderr, "out of memory\n"); return EXIT_FAILURE; }
        size_t got = fread(img + size, 1, cap - size, in);
        if (got == 0) break;
        size += got;
    }
    fclose(in);

    int n_gpus = 1;                      /* AGX_NUM_GPUS=<n> binds more (0 = every visible GPU) */
    const char *env = getenv("AGX_NUM_GPUS");
    if (env) n_gpus = atoi(env);
    if (agx_init(n_gpus) != AGX_OK) {
        fprintf(stderr, "Error: %s\n", agx_last_error());
        return EXIT_FAILURE;
    }
    env = getenv("AGX_PAIRHMM_GATK");   /* 1: mismatch prior Qr/3; 3: that plus the base-quality floor 6 (not the reference) */
    if (env && atoi(env)) agx_pairhmm_set_gatk_mode(atoi(env));
    env = getenv("AGX_PAIRHMM_FP64");   /* every pair through the exact-order FP64 kernel */
    if (env && atoi(env)) agx_pairhmm_set_force_fp64(1);

    const char *err = NULL;
    const double *lh = NULL;
    const int32_t *batch_pairs = NULL;
    int64_t n_out = 0, n_batches = 0;
    int32_t incomplete = 0;
    double *lh_host = NULL;          /* host-split path only */
    int32_t *bp_host = NULL;
    /* the whole file image goes to the library: batch walk and field splitting on the GPU(s); with several GPUs
       bound (AGX_NUM_GPUS > 1) the image is cut into one range of whole batches per GPU */
    int rc = pairhmm_forward_file_image(img, (int64_t)size, &lh, &n_out, &batch_pairs, &n_batches, &incomplete);
    if (rc == AGX_OK) {
        if (incomplete == 1) err = "Error reading reads.\n";
        if (incomplete == 2) err = "Error reading haplotypes.\n";
    } else if (rc == AGX_ERANGE && strstr(agx_last_error(), "line buffer")) {
        rc = split_on_host(img, size, &lh_host, &n_out, &bp_host, &n_batches, &err);
        lh = lh_host;
        batch_pairs = bp_host;
    }
    if (rc != AGX_OK) {
        fprintf(stderr, "Error: code: %d, reason: %s\n", rc, agx_last_error());
        return EXIT_FAILURE;
    }
    int64_t o = 0;
    int iteration = 1;
    for (int64_t b = 0; b < n_batches; b++, iteration++) {
        printf("#batch: %d\n", iteration);
        int64_t cnt = batch_pairs[b];
        for (int64_t k = 0; k < cnt; k++, o++) {
            printf("%f\n", lh[o]);
            fprintf(out, "%f\n", lh[o]);
        }
    }
    printf("#batch: %d\n", iteration);
    if (err) {
        fputs(err, stderr);
        fclose(out);
        return EXIT_FAILURE;
    }
    fclose(out);
    agx_shutdown();
    free(lh_host);
    free(bp_host);
    free(img);
    return EXIT_SUCCESS;
}
