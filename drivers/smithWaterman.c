/*
 * smithWaterman -- drop-in for the reference's Smith-Waterman programs, backed by libagx.so.
 *
 *   usage:  smithWaterman <file_path>            (antidiagonalSmithWaterman.c:190-193)
 *
 * Keeps the reference's command line, input format, stdout lines and exit codes:
 *   - line 1 = atoi() -> number of LINES to consume, printed as "line_num: %d" (:209-210);
 *   - sequences are successive fgets() chunks of a MAX_LINE_LENGTH (1000) byte buffer, trailing
 *     '\n' kept (:201-202, :219-244); a line of >= 999 characters is therefore split, exactly as
 *     the reference does.  AGX_SW_LINE_BUF=<n> raises the buffer (the CUDA twin uses 10000,
 *     smithWaterman.cu:40);
 *   - one "Score: %d" per pair in file order (:348), the dangling first line of an incomplete last
 *     pair echoed (:223-227), then "elapsed %f" in seconds (:351-352).
 * The file image is handed to sw_score_file_image(): the fgets() chunking and the DP (:246-347) both
 * run on the GPU; the host only reads the file and prints.  AGX_NUM_GPUS=<n> binds n GPUs (default 1,
 * 0 = all visible): a large file is then cut into one byte range per GPU.  There is no CPU fallback.
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
    /* slurp the file: the library takes the file image plus (offset, length) of every chunk */
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

    /* header = atoi(first fgets() chunk) (:205-210); printed before any GPU work, like the reference */
    {
        size_t lim = size < (size_t)line_buf - 1 ? size : (size_t)line_buf - 1;
        unsigned char *nl = memchr(img, '\n', lim);
        size_t l = nl ? (size_t)(nl - img) + 1 : lim;
        char tmp[32];
        size_t c = l < sizeof tmp - 1 ? l : sizeof tmp - 1;
        memcpy(tmp, img, c);
        tmp[c] = 0;
        printf("line_num: %d\n", atoi(tmp));
    }
    double iStart = seconds();

    int n_gpus = 1;                               /* AGX_NUM_GPUS=<n> binds more (0 = every visible GPU) */
    env = getenv("AGX_NUM_GPUS");
    if (env) n_gpus = atoi(env);
    int64_t score_cap = (int64_t)(size / 2 + 2), n_pairs = 0, dangling_off = -1;
    int32_t header = 0, dangling_len = 0;
    int32_t *scores = malloc((size_t)score_cap * sizeof *scores);
    if (!scores) { fprintf(stderr, "out of memory\n"); return 1; }
    /* the fgets() chunking (:216-227) and the DP (:246-347) both run on the GPU; the reference's
       scoring constants are :40-43 */
    int rc = agx_init(n_gpus);
    if (rc == AGX_OK)
        rc = sw_score_file_image(img, (int64_t)size, (int32_t)line_buf, 1, -1, -3, -1, scores, score_cap, &n_pairs,
                                 &header, &dangling_off, &dangling_len);
    if (rc != AGX_OK) {
        fprintf(stderr, "Error: code: %d, reason: %s\n", rc, agx_last_error());
        exit(1);
    }
    for (int64_t p = 0; p < n_pairs; p++) printf("Score: %d\n", scores[p]);
    if (dangling_off >= 0) fwrite(img + dangling_off, 1, (size_t)dangling_len, stdout);   /* :223-227 */
    free(scores);

    double iElaps = seconds() - iStart;
    printf("elapsed %f\n", iElaps);
    agx_shutdown();
    free(img);
    return 0;
}
