/*
 * smithWatermanGpu -- the command line of the reference's GPU programs (smithWaterman/hipvers.cpp and
 * smithWaterman.cu), backed by libagx.so, so the published block-size sweep (hiprun.sh) can be re-run
 * shape for shape on a B200.
 *
 *   usage:  smithWatermanGpu <input_file_path> <output_file_path> <block_size>      (hipvers.cpp:392-395)
 *
 * Kept from the reference:
 *   - stdout: "[main] Using Device %d: %s" (:389), "num_of_sequences: %d" (:409), "[main] block_size: %d",
 *     "[main] grid_size: %d" (:474-475), "elapsed %f" (:483);
 *   - line 1 = number of SEQUENCE lines, num_of_sequences / 2 results (:408, :413), 10000-byte fgets()
 *     buffer (:40), the trailing '\n' of a line is a symbol (strlen, :432-444);
 *   - results are APPENDED to <output_file_path> as "Score: %d" lines (:486-495).
 * Different on purpose: <block_size> is echoed but does not steer the launch (libagx picks its own launch
 * geometry per length class); the reference prints uninitialised memory for pairs the file does not
 * contain, this program prints only the pairs it scored; the device is the first one libagx binds
 * (AGX_NUM_GPUS), not the hard-coded `dev = 1`.  There is no CPU fallback.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "agx.h"

#define MAX_LINE_LENGTH 10000 /* hipvers.cpp:40 */

static double seconds(void)
{
    struct timeval tp;
    gettimeofday(&tp, NULL);
    return (double)tp.tv_sec + (double)tp.tv_usec * 1.e-6;
}

int main(int argc, char *argv[])
{
    if (agx_init(getenv("AGX_NUM_GPUS") ? atoi(getenv("AGX_NUM_GPUS")) : 1) != AGX_OK) {
        fprintf(stderr, "Error: %s\n", agx_last_error());
        exit(1);
    }
    printf("[main] Using Device %d: %s\n", agx_device_ordinal(0), agx_device_name(0));
    if (argc != 4) {
        fprintf(stderr, "Usage: %s <input_file_path> <output_file_path> <block_size>\n", argv[0]);
        return 1;
    }
    const char *output_file_path = argv[2];
    int block_size = atoi(argv[3]);
    FILE *input_file = fopen(argv[1], "rb");
    if (input_file == NULL) {
        perror("Error opening file");
        exit(EXIT_FAILURE);
    }
    size_t cap = 1 << 20, size = 0;
    unsigned char *img = malloc(cap);
    for (;;) {
        if (size == cap) { cap *= 2; img = realloc(img, cap); }
        if (!img) { fprintf(stderr, "out of memory\n"); return 1; }
        size_t got = fread(img + size, 1, cap - size, input_file);
        if (got == 0) break;
        size += got;
    }
    fclose(input_file);
    if (size == 0) {
        printf("file is empty");
        return 1;
    }
    /* num_of_sequences = atoi(first fgets() chunk) (:403-408) */
    int num_of_sequences;
    {
        size_t lim = size < MAX_LINE_LENGTH - 1 ? size : MAX_LINE_LENGTH - 1;
        unsigned char *nl = memchr(img, '\n', lim);
        size_t l = nl ? (size_t)(nl - img) + 1 : lim;
        char tmp[32];
        size_t c = l < sizeof tmp - 1 ? l : sizeof tmp - 1;
        memcpy(tmp, img, c);
        tmp[c] = 0;
        num_of_sequences = atoi(tmp);
    }
    printf("num_of_sequences: %d\n", num_of_sequences);
    printf("[main] block_size: %d\n", block_size);
    printf("[main] grid_size: %d\n", num_of_sequences / 2);

    int64_t score_cap = (int64_t)(size / 2 + 2), n_pairs = 0, dangling_off = -1;
    int32_t header = 0, dangling_len = 0;
    int32_t *scores = malloc((size_t)score_cap * sizeof *scores);
    if (!scores) { fprintf(stderr, "out of memory\n"); return 1; }
    double iStart = seconds();
    int rc = sw_score_file_image(img, (int64_t)size, MAX_LINE_LENGTH, 1, -1, -3, -1, scores, score_cap, &n_pairs,
                                 &header, &dangling_off, &dangling_len);     /* constants: hipvers.cpp:41-44 */
    if (rc != AGX_OK) {
        fprintf(stderr, "Error: code: %d, reason: %s\n", rc, agx_last_error());
        exit(1);
    }
    double iElaps = seconds() - iStart;
    printf("elapsed %f\n", iElaps);

    FILE *output_file = fopen(output_file_path, "a");
    if (output_file == NULL) {
        perror("Error opening file");
        return 1;
    }
    /* result_len = num_of_sequences / 2 (:413); never more than the file held */
    int64_t result_len = num_of_sequences > 0 ? num_of_sequences / 2 : 0;
    if (result_len > n_pairs) result_len = n_pairs;
    for (int64_t i = 0; i < result_len; i++) fprintf(output_file, "Score: %d\n", scores[i]);
    if (fclose(output_file) != 0) {
        perror("Error closing file");
        return 1;
    }
    free(scores);
    free(img);
    agx_shutdown();
    return 0;
}
