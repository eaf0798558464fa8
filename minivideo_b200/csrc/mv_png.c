/*
 * mv_png.c -- RGB24 -> PNG, producing the file stb_image_write v1.01 produces for the same pixels.
 *
 * The reference exports PNG through stbi_write_png(name, w, h, 3, rgb, 3*w) (minivideo/src/export.c:532-539;
 * stb_image_write.h is vendored in minivideo/src, v1.01) and, in the default build (no libjpeg), also when
 * 'jpg' was asked for (export.c:652-657).  A PNG is not canonical: the bytes depend on the encoder's row-filter
 * choice and on the token sequence its match finder emits.  To be a byte-for-byte drop-in this file restates
 * those two decisions (stb_image_write.h:848-887 for the filters, :732-790 for the match finder) in its own
 * data structures; everything else (fixed-Huffman deflate bit stream, Adler-32, CRC-32, chunk layout) is the
 * PNG / zlib standard.
 *
 * Row filter: for every row the five filters None/Sub/Up/Average/Paeth are tried in that order (first row:
 * the row above counts as zero) and the first with the smallest sum of |signed byte| wins.
 *
 * Match finder: 16384 buckets keyed by a hash of three bytes; a bucket keeps at most 16 positions and drops
 * its older 8 when full; positions covered by an emitted match are never inserted.  At position i every
 * bucket entry less than 32768 bytes back is measured (<= 258 bytes); the longest wins, ties go to the most
 * recently inserted entry, matches shorter than 3 are ignored.  One step of lazy evaluation: if any entry in
 * the bucket of position i+1 (less than 32767 back) matches longer from i+1, position i becomes a literal.
 */
#include <stdint.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "mv_png.h"

enum { PNG_BUCKETS = 16384, PNG_BUCKET_CAP = 16, PNG_BUCKET_KEEP = 8, PNG_WINDOW = 32768, PNG_MAX_MATCH = 258 };

/* ---- deflate bit stream (LSB first), fixed Huffman codes ------------------------------------------- */

typedef struct { uint8_t *p; size_t n, cap; uint64_t acc; int bits; int failed; } bitsink;

static void sink_reserve(bitsink *s, size_t extra)
{
    if (s->n + extra <= s->cap || s->failed) return;
    size_t cap = s->cap * 2 + extra + 4096;
    uint8_t *q = realloc(s->p, cap);
    if (!q) { s->failed = 1; return; }
    s->p = q; s->cap = cap;
}

static inline void sink_bits(bitsink *s, uint32_t value, int count)
{
    s->acc |= (uint64_t)value << s->bits;
    s->bits += count;
    if (s->bits >= 32) {
        sink_reserve(s, 8);
        if (s->failed) { s->acc = 0; s->bits = 0; return; }
        memcpy(s->p + s->n, &s->acc, 4);        /* little endian host (checked in minivideo_endianness) */
        s->n += 4; s->acc >>= 32; s->bits -= 32;
    }
}

static void sink_finish(bitsink *s)
{
    sink_reserve(s, 16);
    if (s->failed) return;
    while (s->bits > 0) { s->p[s->n++] = (uint8_t)s->acc; s->acc >>= 8; s->bits -= 8; }
    s->bits = 0; s->acc = 0;
}

static uint32_t reverse_bits(uint32_t v, int n)
{
    uint32_t r = 0;
    for (int k = 0; k < n; k++) r |= ((v >> k) & 1u) << (n - 1 - k);
    return r;
}

/* code tables built once: literal/length symbols 0..287 (RFC 1951 3.2.6), length and distance bases */
static uint16_t lit_code[288]; static uint8_t lit_len[288];
static uint16_t len_sym[PNG_MAX_MATCH + 1]; static uint8_t len_xbits[PNG_MAX_MATCH + 1]; static uint16_t len_xval[PNG_MAX_MATCH + 1];
static uint8_t dist_code_rev[30]; static uint16_t dist_base[31]; static uint8_t dist_xbits[30];
static pthread_once_t tables_once = PTHREAD_ONCE_INIT;     /* several writer threads encode at once */
static uint32_t crc_table[256];

static void build_tables(void)
{
    for (uint32_t k = 0; k < 256; k++) {
        uint32_t c = k;
        for (int b = 0; b < 8; b++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
        crc_table[k] = c;
    }
    for (int s = 0; s < 288; s++) {
        int code, n;
        if (s < 144) { code = 0x30 + s; n = 8; }
        else if (s < 256) { code = 0x190 + (s - 144); n = 9; }
        else if (s < 280) { code = s - 256; n = 7; }
        else { code = 0xc0 + (s - 280); n = 8; }
        lit_code[s] = (uint16_t)reverse_bits((uint32_t)code, n); lit_len[s] = (uint8_t)n;
    }
    /* lengths 3..258: eight codes without extra bits, then groups of four with 1..5 extra bits, 258 alone */
    int base = 3, sym = 257;
    for (int group = 0; group < 6; group++) {
        int xb = group, count = group == 0 ? 8 : 4;
        for (int c = 0; c < count && sym < 285; c++, sym++) {
            for (int k = 0; k < (1 << xb) && base + k < PNG_MAX_MATCH; k++) {
                len_sym[base + k] = (uint16_t)sym; len_xbits[base + k] = (uint8_t)xb; len_xval[base + k] = (uint16_t)k;
            }
            base += 1 << xb;
        }
    }
    len_sym[PNG_MAX_MATCH] = 285; len_xbits[PNG_MAX_MATCH] = 0; len_xval[PNG_MAX_MATCH] = 0;
    /* distances 1..32768: codes 0..3 without extra bits, then pairs with 1..13 extra bits */
    int d = 1;
    for (int c = 0; c < 30; c++) {
        int xb = c < 4 ? 0 : (c - 2) / 2;
        dist_base[c] = (uint16_t)d; dist_xbits[c] = (uint8_t)xb; dist_code_rev[c] = (uint8_t)reverse_bits((uint32_t)c, 5);
        d += 1 << xb;
    }
    dist_base[30] = 32769;
}

static inline void put_symbol(bitsink *s, int sym) { sink_bits(s, lit_code[sym], lit_len[sym]); }

static inline void put_match(bitsink *s, int len, int dist)
{
    put_symbol(s, len_sym[len]);
    if (len_xbits[len]) sink_bits(s, len_xval[len], len_xbits[len]);
    int c = 0;
    while (c < 29 && dist >= dist_base[c + 1]) c++;
    sink_bits(s, dist_code_rev[c], 5);
    if (dist_xbits[c]) sink_bits(s, (uint32_t)(dist - dist_base[c]), dist_xbits[c]);
}

/* ---- match finder ----------------------------------------------------------------------------------- */

static inline uint32_t hash3(const uint8_t *p)
{
    uint32_t h = (uint32_t)p[0] + ((uint32_t)p[1] << 8) + ((uint32_t)p[2] << 16);
    h ^= h << 3; h += h >> 5; h ^= h << 4; h += h >> 17; h ^= h << 25; h += h >> 6;
    return h & (PNG_BUCKETS - 1);
}

/* number of equal leading bytes of a[] and b[], at most `limit` */
static inline int common_prefix(const uint8_t *a, const uint8_t *b, int limit)
{
    int n = 0;
    while (n + 8 <= limit) {
        uint64_t x, y;
        memcpy(&x, a + n, 8); memcpy(&y, b + n, 8);
        if (x != y) return n + (__builtin_ctzll(x ^ y) >> 3);
        n += 8;
    }
    while (n < limit && a[n] == b[n]) n++;
    return n;
}

typedef struct { int32_t pos[PNG_BUCKETS][PNG_BUCKET_CAP]; uint8_t count[PNG_BUCKETS]; } buckets_t;

static void deflate_fixed(bitsink *s, const uint8_t *data, int n)
{
    buckets_t *bk = malloc(sizeof *bk);
    if (!bk) { s->failed = 1; return; }
    memset(bk->count, 0, sizeof bk->count);
    sink_bits(s, 1, 1);                 /* last block */
    sink_bits(s, 1, 2);                 /* fixed Huffman */
    int i = 0;
    while (i < n - 3) {
        const uint32_t h = hash3(data + i);
        const int room = n - i < PNG_MAX_MATCH ? n - i : PNG_MAX_MATCH;
        int best = 3, from = -1;
        int32_t *e = bk->pos[h];
        int cnt = bk->count[h];
        for (int j = 0; j < cnt; j++) {
            const int p = e[j];
            if (p <= i - PNG_WINDOW) continue;
            /* only a match of at least `best` bytes can replace the current one: look at its last byte first */
            if (best <= room && data[p + best - 1] != data[i + best - 1]) continue;
            const int d = common_prefix(data + p, data + i, room);
            if (d >= best) { best = d; from = p; }
        }
        if (cnt == PNG_BUCKET_CAP) {
            memmove(e, e + PNG_BUCKET_KEEP, sizeof(int32_t) * PNG_BUCKET_KEEP);
            cnt = PNG_BUCKET_KEEP;
        }
        e[cnt++] = i;
        bk->count[h] = (uint8_t)cnt;

        if (from >= 0) {                /* would the next position do better? */
            const uint32_t h1 = hash3(data + i + 1);
            const int room1 = n - i - 1 < PNG_MAX_MATCH ? n - i - 1 : PNG_MAX_MATCH;
            const int32_t *e1 = bk->pos[h1];
            const int cnt1 = bk->count[h1];
            if (best < room1) {
                for (int j = 0; j < cnt1; j++) {
                    const int p = e1[j];
                    if (p <= i - (PNG_WINDOW - 1)) continue;
                    if (data[p + best] != data[i + 1 + best]) continue;
                    if (common_prefix(data + p, data + i + 1, room1) > best) { from = -1; break; }
                }
            }
        }
        if (from >= 0) { put_match(s, best, i - from); i += best; }
        else { put_symbol(s, data[i]); i++; }
    }
    for (; i < n; i++) put_symbol(s, data[i]);
    put_symbol(s, 256);
    sink_finish(s);
    free(bk);
}

static uint32_t adler32(const uint8_t *d, size_t n)
{
    uint32_t a = 1, b = 0;
    while (n) {
        size_t k = n < 5552 ? n : 5552;
        for (size_t i = 0; i < k; i++) { a += d[i]; b += a; }
        a %= 65521; b %= 65521; d += k; n -= k;
    }
    return (b << 16) | a;
}

static uint32_t crc32_png(const uint8_t *d, size_t n)       /* tables: build_tables(), once */
{
    uint32_t c = ~0u;
    for (size_t i = 0; i < n; i++) c = (c >> 8) ^ crc_table[(c ^ d[i]) & 255];
    return ~c;
}

/* ---- row filters ------------------------------------------------------------------------------------ */

static inline int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    return pb <= pc ? b : c;
}

static inline long cost(const uint8_t *v, int n)
{
    long c = 0;
    for (int i = 0; i < n; i++) c += abs((int8_t)v[i]);
    return c;
}

/* writes filter byte + filtered row into dst[0..row_bytes]; `up` is the row above or a row of zeros */
static void filter_row(uint8_t *dst, const uint8_t *cur, const uint8_t *up, int row_bytes, int bpp, uint8_t *scratch)
{
    uint8_t *cand[5];
    for (int k = 0; k < 5; k++) cand[k] = scratch + (size_t)k * row_bytes;
    memcpy(cand[0], cur, (size_t)row_bytes);
    for (int i = 0; i < row_bytes; i++) {
        const int a = i >= bpp ? cur[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
        cand[1][i] = (uint8_t)(cur[i] - a);
        cand[2][i] = (uint8_t)(cur[i] - b);
        cand[3][i] = (uint8_t)(cur[i] - ((a + b) >> 1));
        cand[4][i] = (uint8_t)(cur[i] - paeth(a, b, c));
    }
    int pick = 0; long low = cost(cand[0], row_bytes);
    for (int k = 1; k < 5; k++) { long c = cost(cand[k], row_bytes); if (c < low) { low = c; pick = k; } }
    dst[0] = (uint8_t)pick;
    memcpy(dst + 1, cand[pick], (size_t)row_bytes);
}

static uint8_t *put_be32(uint8_t *o, uint32_t v) { o[0] = (uint8_t)(v >> 24); o[1] = (uint8_t)(v >> 16); o[2] = (uint8_t)(v >> 8); o[3] = (uint8_t)v; return o + 4; }

uint8_t *mvt_png_encode(const uint8_t *rgb, int w, int h, size_t *out_len)
{
    if (out_len) *out_len = 0;
    if (!rgb || !out_len || w < 1 || h < 1 || (long long)w * 3 + 1 > 0x7fffffff / h) return NULL;
    pthread_once(&tables_once, build_tables);
    const int row_bytes = 3 * w;
    const int n = (row_bytes + 1) * h;
    uint8_t *filt = malloc((size_t)n), *scratch = malloc((size_t)row_bytes * 6);
    if (!filt || !scratch) { free(filt); free(scratch); return NULL; }
    uint8_t *zero = scratch + (size_t)row_bytes * 5;
    memset(zero, 0, (size_t)row_bytes);
    for (int y = 0; y < h; y++)
        filter_row(filt + (size_t)y * (row_bytes + 1), rgb + (size_t)y * row_bytes,
                   y ? rgb + (size_t)(y - 1) * row_bytes : zero, row_bytes, 3, scratch);
    free(scratch);

    bitsink s = {0};
    sink_reserve(&s, (size_t)n / 2 + 64);
    if (!s.failed) { s.p[s.n++] = 0x78; s.p[s.n++] = 0x5e; }    /* 32 K window, "fast" level hint */
    deflate_fixed(&s, filt, n);
    const uint32_t ad = adler32(filt, (size_t)n);
    free(filt);
    if (s.failed) { free(s.p); return NULL; }
    uint8_t tail[4]; put_be32(tail, ad);
    sink_reserve(&s, 4);
    if (s.failed) { free(s.p); return NULL; }
    memcpy(s.p + s.n, tail, 4); s.n += 4;

    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    const size_t total = 8 + (12 + 13) + (12 + s.n) + 12;
    uint8_t *png = malloc(total), *o = png;
    if (!png) { free(s.p); return NULL; }
    memcpy(o, sig, 8); o += 8;
    o = put_be32(o, 13); memcpy(o, "IHDR", 4); o += 4;
    o = put_be32(o, (uint32_t)w); o = put_be32(o, (uint32_t)h);
    *o++ = 8; *o++ = 2; *o++ = 0; *o++ = 0; *o++ = 0;           /* 8 bits, truecolour, deflate, adaptive, no interlace */
    o = put_be32(o, crc32_png(o - 17, 17));
    o = put_be32(o, (uint32_t)s.n); memcpy(o, "IDAT", 4); o += 4;
    memcpy(o, s.p, s.n); o += s.n;
    o = put_be32(o, crc32_png(o - s.n - 4, s.n + 4));
    free(s.p);
    o = put_be32(o, 0); memcpy(o, "IEND", 4); o += 4;
    o = put_be32(o, crc32_png(o - 4, 4));
    *out_len = (size_t)(o - png);
    return png;
}
