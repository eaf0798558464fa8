/* Test helper (built by tests/test_front_fuzz.py with -fsanitize=address,undefined): parse many corrupted
 * copies of a valid stream; the front end must answer SUCCESS / FAILURE / UNSUPPORTED and never touch memory
 * it does not own. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mvfront.h"

static unsigned long long rng = 0x9E3779B97F4A7C15ull;
static unsigned rnd(void) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (unsigned)(rng >> 32); }

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t *orig = malloc((size_t)n), *buf = malloc((size_t)n);
    if (fread(orig, 1, (size_t)n, f) != (size_t)n) return 2;
    fclose(f);
    const int rounds = atoi(argv[2]);
    int opened = 0, parsed = 0;
    for (int r = 0; r < rounds; r++) {
        memcpy(buf, orig, (size_t)n);
        size_t len = (size_t)n;
        const int kind = r % 4;
        if (kind == 0) for (int k = 0; k < 1 + (int)(rnd() % 8); k++) buf[rnd() % len] ^= (uint8_t)(1u << (rnd() % 8));   /* bit flips */
        else if (kind == 1) len = 1 + rnd() % len;                                                                        /* truncation */
        else if (kind == 2) { size_t a = rnd() % len, c = 1 + rnd() % 64; if (a + c > len) c = len - a; for (size_t i = 0; i < c; i++) buf[a + i] = (uint8_t)rnd(); }
        else { size_t a = rnd() % len, c = 1 + rnd() % 256; if (a + c > len) c = len - a; memset(buf + a, rnd() & 1 ? 0 : 0xff, c); }
        uint8_t *exact = malloc(len ? len : 1);          /* exact-size copy: the sanitizer sees any read past the stream */
        memcpy(exact, buf, len);
        mvf_stream *s = NULL;
        if (mvf_open_annexb(exact, len, &s) == 1) {
            opened++;
            mvf_info in;
            mvf_get_info(s, &in);
            if (in.n_idr > 0 && in.width_mbs > 0 && in.width_mbs <= 64 && in.height_mbs <= 64) {
                const size_t N = (size_t)in.width_mbs * in.height_mbs, P = (size_t)in.n_idr;
                mvf_batch b;
                memset(&b, 0, sizeof b);
                int32_t *status = malloc(P * sizeof *status);
                b.mb_kind = malloc(N * P); b.i16_mode = malloc(N * P); b.chroma_mode = malloc(N * P); b.qp_y = malloc(N * P);
                b.cbp = malloc(N * P); b.luma_modes = malloc(N * P * 16); b.coeff = malloc(N * P * 768);
                if (mvf_parse_pictures(s, NULL, 0, (int)P, &b, 1 + r % 3) == 1) parsed++;
                b.status = status;                      /* tolerant mode: per-picture codes, the call itself succeeds */
                mvf_parse_pictures(s, NULL, 0, (int)P, &b, 1 + r % 3);
                mvf_packed_batch pk;
                memset(&pk, 0, sizeof pk);
                pk.mb_kind = b.mb_kind; pk.i16_mode = b.i16_mode; pk.chroma_mode = b.chroma_mode; pk.qp_y = b.qp_y; pk.luma_modes = b.luma_modes;
                pk.nz_blocks = malloc(N * P * 4); pk.word_off = malloc(N * P * 4); pk.pic_off = malloc((P + 1) * 8);
                pk.words_capacity = N * P * 408; pk.words = malloc(pk.words_capacity * 2);
                mvf_parse_pictures_packed(s, NULL, 0, (int)P, &pk, 2);
                pk.status = status;
                mvf_parse_pictures_packed(s, NULL, 0, (int)P, &pk, 2);
                for (int g = 0; g < mvf_generation_count(s); g++) { mvf_info gi; mvf_get_generation_info(s, g, &gi); }
                free(status);
                int32_t *sel = malloc(sizeof(int32_t) * (P + 8));
                mvf_select_idr(s, 3, r % 3, sel);
                free(sel); free(pk.nz_blocks); free(pk.word_off); free(pk.pic_off); free(pk.words);
                free(b.mb_kind); free(b.i16_mode); free(b.chroma_mode); free(b.qp_y); free(b.cbp); free(b.luma_modes); free(b.coeff);
            }
            mvf_close(s);
        }
        free(exact);
    }
    free(orig); free(buf);
    printf("rounds=%d opened=%d parsed=%d\n", rounds, opened, parsed);
    return 0;
}
