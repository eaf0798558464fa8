/* Development aid: single-thread (or -t N) parse speed of the host front end on a stream file.
 *   gcc -O2 -g -Iinclude -Iminivideo_b200/csrc -o /tmp/front_bench tests/tools/front_bench.c minivideo_b200/csrc/h264_front.c -lm -lpthread
 *   /tmp/front_bench stream.264 [reps] [threads] */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "mvfront.h"

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t *data = malloc((size_t)n);
    if (fread(data, 1, (size_t)n, f) != (size_t)n) return 2;
    fclose(f);
    const int reps = argc > 2 ? atoi(argv[2]) : 5, threads = argc > 3 ? atoi(argv[3]) : 1;
    mvf_stream *s = NULL;
    if (mvf_open_annexb(data, (size_t)n, &s) != 1) { fprintf(stderr, "%s\n", mvf_last_error(NULL)); return 1; }
    mvf_info in;
    mvf_get_info(s, &in);
    const size_t N = (size_t)in.width_mbs * in.height_mbs, P = (size_t)in.n_idr;
    mvf_packed_batch pk;
    memset(&pk, 0, sizeof pk);
    pk.mb_kind = malloc(N * P); pk.i16_mode = malloc(N * P); pk.chroma_mode = malloc(N * P); pk.qp_y = malloc(N * P);
    pk.luma_modes = malloc(N * P * 16); pk.nz_blocks = malloc(N * P * 4); pk.word_off = malloc(N * P * 4);
    pk.pic_off = malloc((P + 1) * 8); pk.words_capacity = N * P * 408; pk.words = malloc(pk.words_capacity * 2);
    mvf_parser *ps = NULL;
    mvf_parser_create(s, threads, &ps);
    mvf_parser_parse_packed(ps, NULL, 0, (int)P, &pk);
    double best = 1e30, sum = 0;
    for (int r = 0; r < reps; r++) {
        const double t0 = now();
        if (mvf_parser_parse_packed(ps, NULL, 0, (int)P, &pk) != 1) { fprintf(stderr, "%s\n", mvf_parser_last_error(ps)); return 1; }
        const double dt = now() - t0;
        sum += dt;
        if (dt < best) best = dt;
    }
    printf("%zu pictures %dx%d MBs, %d thread(s): best %.1f pictures/s (%.0f ns per macroblock), mean %.1f pictures/s, %llu words per picture\n",
           P, in.width_mbs, in.height_mbs, threads, P / best, 1e9 * best / (P * N), reps * P / sum, (unsigned long long)(pk.pic_off[P] / P));
    mvf_parser_destroy(ps);
    mvf_close(s);
    return 0;
}
