/*
 * h264_synth.c -- minimal CAVLC intra-only H.264 encoder with random modes and
 * residuals (see include/mvsynth.h).  Written from the H.264 specification
 * (clauses 7.3, 9.1, 9.2); the constraints it honours come from the reference
 * decoder's parser (citations: SURVEY.md section 8(c)).
 *
 * The encoder never reconstructs samples.  For every macroblock it draws
 *   mb kind, prediction modes (only modes whose neighbour samples exist),
 *   coded_block_pattern, mb_qp_delta, transform coefficient levels
 * writes them with ue/se/me/CAVLC, and mirrors them into the mvgpu.h SoA.
 */
#include "mvsynth.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* RNG: splitmix64-seeded xorshift64*                                        */

typedef struct { uint64_t s; } rng_t;

static uint64_t rng_next(rng_t *r)
{
    uint64_t x = r->s;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    r->s = x;
    return x * 0x2545F4914F6CDD1DULL;
}
static void rng_seed(rng_t *r, uint64_t seed)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    r->s = (z ^ (z >> 31)) | 1;
}
static uint32_t rng_below(rng_t *r, uint32_t n) { return (uint32_t)((rng_next(r) >> 32) % n); }
static double   rng_unit(rng_t *r) { return ((rng_next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

/* ------------------------------------------------------------------------ */
/* RBSP bit writer + NAL emission with emulation prevention (7.4.1)          */

typedef struct {
    uint8_t *buf; size_t cap, len;
    uint32_t acc; int nacc;
    int overflow;
} bw_t;

static void bw_byte(bw_t *b, uint8_t v)
{
    if (b->len < b->cap) b->buf[b->len++] = v; else b->overflow = 1;
}
static void bw_put(bw_t *b, int n, uint32_t v)
{
    for (int i = n - 1; i >= 0; i--) {
        b->acc = (b->acc << 1) | ((v >> i) & 1);
        if (++b->nacc == 8) { bw_byte(b, (uint8_t)b->acc); b->acc = 0; b->nacc = 0; }
    }
}
static void bw_ue(bw_t *b, uint32_t v)
{
    uint32_t x = v + 1; int n = 0;
    while ((x >> n) > 1) n++;
    bw_put(b, n, 0);
    bw_put(b, n + 1, x);
}
static void bw_se(bw_t *b, int v) { bw_ue(b, v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v)); }
static void bw_trailing(bw_t *b)
{
    bw_put(b, 1, 1);
    while (b->nacc) bw_put(b, 1, 0);
}

typedef struct { uint8_t *buf; size_t cap, len; int overflow; } sink_t;

static void sink_byte(sink_t *s, uint8_t v)
{
    if (s->len < s->cap) s->buf[s->len++] = v; else s->overflow = 1;
}
static void emit_nal(sink_t *s, uint8_t header, const bw_t *rbsp)
{
    sink_byte(s, 0); sink_byte(s, 0); sink_byte(s, 0); sink_byte(s, 1);
    sink_byte(s, header);
    int zeros = 0;
    for (size_t i = 0; i < rbsp->len; i++) {
        uint8_t v = rbsp->buf[i];
        if (zeros >= 2 && v <= 3) { sink_byte(s, 3); zeros = 0; }
        sink_byte(s, v);
        zeros = (v == 0) ? zeros + 1 : 0;
    }
}

#include "h264_cavlc_tables.h"

static void bw_str(bw_t *b, const char *s)
{
    for (; *s; s++) bw_put(b, 1, *s == '1');
}

/* frame zig-zag scans generated from the definition (8.5.6, 8.5.7):
 * zz[k] = row*n + col of the k-th coefficient */
static void make_zigzag(int n, uint8_t *zz)
{
    int r = 0, c = 0, up = 1;   /* first move is to the right, then down-left */
    for (int k = 0; k < n * n; k++) {
        zz[k] = (uint8_t)(r * n + c);
        if (up) {               /* moving up-right */
            if (c == n - 1) { r++; up = 0; }
            else if (r == 0) { c++; up = 0; }
            else { r--; c++; }
        } else {                /* moving down-left */
            if (r == n - 1) { c++; up = 1; }
            else if (c == 0) { r++; up = 1; }
            else { r++; c--; }
        }
    }
}

/* position of 4x4 luma block blk inside the MB, in 4x4 units (6.4.3) */
static inline int blk_x(int blk) { return (blk & 1) + 2 * ((blk >> 2) & 1); }
static inline int blk_y(int blk) { return ((blk >> 1) & 1) + 2 * (blk >> 3); }

/* ------------------------------------------------------------------------ */

typedef struct {
    const mvs_params *p;
    rng_t rng;
    int W, H;                 /* in MBs */
    uint8_t *tot_luma;        /* [H*4][W*4] TotalCoeff of each 4x4 luma block   */
    uint8_t *tot_chroma[2];   /* [H*2][W*2]                                     */
    int8_t  *mode_grid;       /* [H*4][W*4] Intra4x4/8x8 mode, 2 for I16x16     */
    uint8_t zz4[16], zz8[64];
    uint8_t cbp_to_codenum[48];
} enc_t;

/* 9.2.1: write coeff_token */
static void put_coeff_token(bw_t *b, int nC, int tc, int t1)
{
    if (nC == -1) { bw_put(b, ctc_len[t1][tc], ctc_code[t1][tc]); return; }
    if (nC >= 8) { bw_put(b, 6, tc == 0 ? 3u : (uint32_t)(((tc - 1) << 2) | t1)); return; }
    int t = nC < 2 ? 0 : (nC < 4 ? 1 : 2);
    bw_put(b, ct_len[t][t1][tc], ct_code[t][t1][tc]);
}

/* 9.2: encode one residual block given in scan order; returns TotalCoeff */
static int put_residual_block(bw_t *b, const int *coef, int max_num, int nC)
{
    int idx[16], tc = 0;
    for (int k = 0; k < max_num; k++) if (coef[k]) idx[tc++] = k;

    int t1 = 0;
    while (t1 < 3 && t1 < tc && abs(coef[idx[tc - 1 - t1]]) == 1) t1++;
    put_coeff_token(b, nC, tc, t1);
    if (tc == 0) return 0;

    int suffix_len = (tc > 10 && t1 < 3) ? 1 : 0;
    for (int i = 0; i < tc; i++) {
        int level = coef[idx[tc - 1 - i]];
        if (i < t1) { bw_put(b, 1, level < 0); continue; }
        int code = level > 0 ? 2 * level - 2 : -2 * level - 1;
        if (i == t1 && t1 < 3) code -= 2;
        if (suffix_len == 0) {
            if (code < 14) { bw_put(b, code, 0); bw_put(b, 1, 1); }
            else if (code < 30) { bw_put(b, 14, 0); bw_put(b, 1, 1); bw_put(b, 4, (uint32_t)(code - 14)); }
            else { bw_put(b, 15, 0); bw_put(b, 1, 1); bw_put(b, 12, (uint32_t)(code - 30)); }
        } else {
            if (code < (15 << suffix_len)) {
                bw_put(b, code >> suffix_len, 0); bw_put(b, 1, 1);
                bw_put(b, suffix_len, (uint32_t)(code & ((1 << suffix_len) - 1)));
            } else {
                bw_put(b, 15, 0); bw_put(b, 1, 1);
                bw_put(b, 12, (uint32_t)(code - (15 << suffix_len)));
            }
        }
        if (suffix_len == 0) suffix_len = 1;
        if (abs(level) > (3 << (suffix_len - 1)) && suffix_len < 6) suffix_len++;
    }

    int total_zeros = idx[tc - 1] + 1 - tc;
    if (tc < max_num) {
        if (max_num == 4) bw_str(b, tz2x2[tc - 1][total_zeros]);
        else              bw_str(b, tz4x4[tc - 1][total_zeros]);
    }
    int zeros_left = total_zeros;
    for (int i = 0; i < tc - 1 && zeros_left > 0; i++) {
        int run = idx[tc - 1 - i] - idx[tc - 2 - i] - 1;
        bw_str(b, runb[(zeros_left > 7 ? 7 : zeros_left) - 1][run]);
        zeros_left -= run;
    }
    return tc;
}

/* draw a random block of n coefficients in scan order, starting at `first` */
static void draw_block(enc_t *e, int *coef, int n, int first, double mean_tc, int max_level)
{
    memset(coef, 0, sizeof(int) * (size_t)n);
    double q = mean_tc / (1.0 + mean_tc);                 /* geometric, mean = mean_tc */
    int tc = (int)floor(log(rng_unit(&e->rng)) / log(q));
    int room = n - first;
    if (tc > room) tc = room;
    double scale = e->p->level_scale_x10 / 10.0;
    for (int i = 0; i < tc; i++) {
        double u = rng_unit(&e->rng);
        int pos = first + (int)(room * u * u);            /* biased to low frequencies */
        while (coef[pos]) pos = first + (pos - first + 1) % room;
        int mag = 1 + (int)floor(-log(rng_unit(&e->rng)) * scale);
        if (mag > max_level) mag = max_level;
        coef[pos] = (rng_next(&e->rng) & 1) ? mag : -mag;
    }
}

static int pick_kind(enc_t *e)
{
    const mvs_params *p = e->p;
    if (p->force_kind >= 0) return p->force_kind;
    int w8 = (p->profile_idc >= 100 && p->transform8x8) ? p->w_i8x8 : 0;
    int tot = p->w_i4x4 + w8 + p->w_i16x16;
    int r = (int)rng_below(&e->rng, (uint32_t)tot);
    if (r < p->w_i4x4) return 0;
    if (r < p->w_i4x4 + w8) return 1;
    return 2;
}

/* a random Intra4x4/8x8 mode legal for the given neighbour availability */
static int pick_nxn_mode(enc_t *e, int left, int up)
{
    int ok[9], n = 0;
    for (int m = 0; m < 9; m++) {
        int need_up = (m == 0 || m == 3 || m == 7), need_left = (m == 1 || m == 8);
        int need_all = (m >= 4 && m <= 6);   /* left, up and up-left */
        if (need_up && !up) continue;
        if (need_left && !left) continue;
        if (need_all && !(left && up)) continue;
        ok[n++] = m;
    }
    if (e->p->force_mode >= 0)
        for (int i = 0; i < n; i++) if (ok[i] == e->p->force_mode) return ok[i];
    return ok[rng_below(&e->rng, (uint32_t)n)];
}
/* Intra16x16: 0 V, 1 H, 2 DC, 3 Plane.  Chroma: 0 DC, 1 H, 2 V, 3 Plane */
static int pick_16_mode(enc_t *e, int left, int up, int chroma)
{
    int ok[4], n = 0;
    for (int m = 0; m < 4; m++) {
        int is_v = chroma ? (m == 2) : (m == 0), is_h = (m == 1);
        int is_dc = chroma ? (m == 0) : (m == 2);
        if (is_v && !up) continue;
        if (is_h && !left) continue;
        if (m == 3 && !(left && up)) continue;
        (void)is_dc;
        ok[n++] = m;
    }
    if (e->p->force_mode >= 0)
        for (int i = 0; i < n; i++) if (ok[i] == (e->p->force_mode & 3)) return ok[i];
    return ok[rng_below(&e->rng, (uint32_t)n)];
}

static int luma_nC(const enc_t *e, int X4, int Y4)
{
    int w4 = e->W * 4;
    int a = X4 > 0, b = Y4 > 0;
    int nA = a ? e->tot_luma[Y4 * w4 + X4 - 1] : 0;
    int nB = b ? e->tot_luma[(Y4 - 1) * w4 + X4] : 0;
    if (a && b) return (nA + nB + 1) >> 1;
    return a ? nA : (b ? nB : 0);
}
static int chroma_nC(const enc_t *e, int c, int X2, int Y2)
{
    int w2 = e->W * 2;
    int a = X2 > 0, b = Y2 > 0;
    int nA = a ? e->tot_chroma[c][Y2 * w2 + X2 - 1] : 0;
    int nB = b ? e->tot_chroma[c][(Y2 - 1) * w2 + X2] : 0;
    if (a && b) return (nA + nB + 1) >> 1;
    return a ? nA : (b ? nB : 0);
}

/* predicted Intra4x4/8x8 mode for the block whose top-left 4x4 cell is (X4,Y4)
 * (8.3.1.1 / 8.3.2.1 with every MB intra and constrained_intra_pred == 0) */
static int predicted_mode(const enc_t *e, int X4, int Y4)
{
    if (X4 == 0 || Y4 == 0) return 2;
    int w4 = e->W * 4;
    int a = e->mode_grid[Y4 * w4 + X4 - 1], b = e->mode_grid[(Y4 - 1) * w4 + X4];
    return a < b ? a : b;
}

static void put_pred_mode(bw_t *b, int mode, int pred)
{
    if (mode == pred) { bw_put(b, 1, 1); return; }
    bw_put(b, 1, 0);
    bw_put(b, 3, (uint32_t)(mode < pred ? mode : mode - 1));
}

typedef struct {
    uint8_t *kind, *i16, *cm, *cbp, *modes; int8_t *qp; int16_t *coeff;
} soa_t;

static void encode_mb(enc_t *e, bw_t *b, int mx, int my, int *qp_prev, soa_t *s, size_t mbi)
{
    const mvs_params *p = e->p;
    int W4 = e->W * 4, W2 = e->W * 2;
    int left = mx > 0, up = my > 0;
    int kind = pick_kind(e);
    int16_t cf[384];
    memset(cf, 0, sizeof cf);
    uint8_t modes[16] = {0};
    int i16_mode = 0;

    /* --- prediction modes ------------------------------------------------- */
    int pred[16];
    if (kind == 0) {
        for (int blk = 0; blk < 16; blk++) {
            int X4 = mx * 4 + blk_x(blk), Y4 = my * 4 + blk_y(blk);
            pred[blk] = predicted_mode(e, X4, Y4);
            modes[blk] = (uint8_t)pick_nxn_mode(e, X4 > 0, Y4 > 0);
            e->mode_grid[Y4 * W4 + X4] = (int8_t)modes[blk];
        }
    } else if (kind == 1) {
        for (int b8 = 0; b8 < 4; b8++) {
            int X4 = mx * 4 + (b8 & 1) * 2, Y4 = my * 4 + (b8 >> 1) * 2;
            pred[b8] = predicted_mode(e, X4, Y4);
            modes[b8] = (uint8_t)pick_nxn_mode(e, X4 > 0, Y4 > 0);
            for (int dy = 0; dy < 2; dy++)
                for (int dx = 0; dx < 2; dx++)
                    e->mode_grid[(Y4 + dy) * W4 + X4 + dx] = (int8_t)modes[b8];
        }
    } else {
        i16_mode = pick_16_mode(e, left, up, 0);
        for (int blk = 0; blk < 16; blk++)
            e->mode_grid[(my * 4 + blk_y(blk)) * W4 + mx * 4 + blk_x(blk)] = 2;
    }
    int chroma_mode = pick_16_mode(e, left, up, 1);

    /* --- coded block pattern ---------------------------------------------- */
    int cbp_l = 0, cbp_c = (int)rng_below(&e->rng, 3);
    if (kind == 2) cbp_l = ((int)rng_below(&e->rng, 100) < p->luma_cbp_percent) ? 15 : 0;
    else for (int i = 0; i < 4; i++)
        if ((int)rng_below(&e->rng, 100) < p->luma_cbp_percent) cbp_l |= 1 << i;

    /* --- QP ---------------------------------------------------------------- */
    int has_residual = (kind == 2) || cbp_l || cbp_c;
    int qp = *qp_prev, delta = 0;
    if (has_residual) {
        for (int tries = 0; tries < 64; tries++) {
            delta = (int)rng_below(&e->rng, 5) - 2;
            int q = *qp_prev + delta;
            if (q < p->qp_min || q > p->qp_max) continue;
            if (kind == 2 && q == 36) continue;   /* reference UB: h264_transform.c:797-808 */
            break;
        }
        qp = *qp_prev + delta;
        if (qp < p->qp_min || qp > p->qp_max || (kind == 2 && qp == 36)) {
            /* deterministic way out: walk to a legal value */
            qp = *qp_prev;
            if (qp < p->qp_min) qp = p->qp_min;
            if (qp > p->qp_max) qp = p->qp_max;
            if (kind == 2 && qp == 36) qp = (qp + 1 <= p->qp_max) ? 37 : 35;
            delta = qp - *qp_prev;
        }
    }

    /* --- syntax: mb_type .. mb_qp_delta ------------------------------------ */
    if (kind == 2) bw_ue(b, (uint32_t)(1 + i16_mode + 4 * cbp_c + (cbp_l ? 12 : 0)));
    else {
        bw_ue(b, 0);                                            /* I_NxN */
        if (p->profile_idc >= 100 && p->transform8x8) bw_put(b, 1, kind == 1);
        int nb = kind == 0 ? 16 : 4;
        for (int i = 0; i < nb; i++) put_pred_mode(b, modes[i], pred[i]);
    }
    bw_ue(b, (uint32_t)chroma_mode);
    if (kind != 2) bw_ue(b, e->cbp_to_codenum[cbp_c * 16 + cbp_l]);
    if (has_residual) bw_se(b, delta);

    /* --- residual ----------------------------------------------------------- */
    double mean4 = p->mean_coeffs_x10 / 10.0;
    int coef[64], sub[16];
    if (has_residual) {
        if (kind == 2) {       /* Intra16x16 DC: one 16-coefficient block, always sent */
            draw_block(e, coef, 16, 0, mean4, 4 * p->max_level);
            put_residual_block(b, coef, 16, luma_nC(e, mx * 4, my * 4));
            for (int k = 0; k < 16; k++) {
                int r = e->zz4[k] >> 2, c = e->zz4[k] & 3;      /* matrix position of level k */
                int blk = (r & 1) * 2 + (r >> 1) * 8 + (c & 1) + (c >> 1) * 4;
                cf[blk * 16] = (int16_t)coef[k];
            }
        }
        for (int b8 = 0; b8 < 4; b8++) {
            int coded = (cbp_l >> b8) & 1;
            if (kind == 1 && coded) draw_block(e, coef, 64, 0, mean4 * 3.0, p->max_level);
            for (int i4 = 0; i4 < 4; i4++) {
                int blk = b8 * 4 + i4;
                int X4 = mx * 4 + blk_x(blk), Y4 = my * 4 + blk_y(blk);
                int tc = 0;
                if (coded) {
                    int nC = luma_nC(e, X4, Y4);
                    if (kind == 0) {
                        draw_block(e, sub, 16, 0, mean4, p->max_level);
                        tc = put_residual_block(b, sub, 16, nC);
                        for (int k = 0; k < 16; k++) cf[blk * 16 + k] = (int16_t)sub[k];
                    } else if (kind == 1) {
                        for (int k = 0; k < 16; k++) sub[k] = coef[4 * k + i4];
                        tc = put_residual_block(b, sub, 16, nC);
                        for (int k = 0; k < 16; k++) cf[b8 * 64 + 4 * k + i4] = (int16_t)sub[k];
                    } else {
                        draw_block(e, sub, 15, 0, mean4, p->max_level);
                        tc = put_residual_block(b, sub, 15, nC);
                        for (int k = 0; k < 15; k++) cf[blk * 16 + 1 + k] = (int16_t)sub[k];
                    }
                }
                e->tot_luma[Y4 * W4 + X4] = (uint8_t)tc;
            }
        }
        /* chroma DC (both planes), then chroma AC (both planes) */
        for (int c = 0; c < 2; c++) {
            if (cbp_c & 3) {
                draw_block(e, sub, 4, 0, 1.5, 2 * p->max_level);
                put_residual_block(b, sub, 4, -1);
                for (int k = 0; k < 4; k++) cf[256 + c * 64 + k * 16] = (int16_t)sub[k];
            }
        }
        for (int c = 0; c < 2; c++)
            for (int blk = 0; blk < 4; blk++) {
                int X2 = mx * 2 + (blk & 1), Y2 = my * 2 + (blk >> 1);
                int tc = 0;
                if (cbp_c & 2) {
                    draw_block(e, sub, 15, 0, mean4 * 0.6, p->max_level);
                    tc = put_residual_block(b, sub, 15, chroma_nC(e, c, X2, Y2));
                    for (int k = 0; k < 15; k++) cf[256 + c * 64 + blk * 16 + 1 + k] = (int16_t)sub[k];
                }
                e->tot_chroma[c][Y2 * W2 + X2] = (uint8_t)tc;
            }
    } else {
        for (int blk = 0; blk < 16; blk++)
            e->tot_luma[(my * 4 + blk_y(blk)) * W4 + mx * 4 + blk_x(blk)] = 0;
        for (int c = 0; c < 2; c++)
            for (int blk = 0; blk < 4; blk++)
                e->tot_chroma[c][(my * 2 + (blk >> 1)) * W2 + mx * 2 + (blk & 1)] = 0;
    }
    *qp_prev = qp;

    /* --- SoA mirror ---------------------------------------------------------- */
    if (s->kind)  s->kind[mbi] = (uint8_t)kind;
    if (s->i16)   s->i16[mbi] = (uint8_t)i16_mode;
    if (s->cm)    s->cm[mbi] = (uint8_t)chroma_mode;
    if (s->cbp)   s->cbp[mbi] = (uint8_t)(cbp_c << 4 | cbp_l);
    if (s->qp)    s->qp[mbi] = (int8_t)qp;
    if (s->modes) memcpy(s->modes + mbi * 16, modes, 16);
    if (s->coeff) memcpy(s->coeff + mbi * 384, cf, sizeof cf);
}

/* 7.3.2.1.1.1 scaling_list(): delta_scale sequence for a list with no zero entry */
static void put_scaling_list(bw_t *b, const uint8_t *list, int n)
{
    int last = 8;
    for (int j = 0; j < n; j++) {
        int d = (int)list[j] - last;
        if (d > 127) d -= 256;
        if (d < -128) d += 256;
        bw_se(b, d);
        last = list[j];
    }
}

void mvs_default_params(mvs_params *p)
{
    memset(p, 0, sizeof *p);
    p->width_mbs = 22; p->height_mbs = 18; p->n_pics = 1;
    p->profile_idc = 66;
    p->w_i4x4 = 1; p->w_i8x8 = 1; p->w_i16x16 = 1;
    p->init_qp = 26; p->qp_min = 20; p->qp_max = 32;
    p->luma_cbp_percent = 50;
    p->mean_coeffs_x10 = 40;
    p->level_scale_x10 = 12;
    p->max_level = 64;
    p->poc_type = 0;
    p->force_mode = -1; p->force_kind = -1;
    p->seed = 0xC0FFEE;
}

size_t mvs_stream_bound(const mvs_params *p)
{
    /* worst case per MB: ~27 blocks x (16 coeffs x ~28 bits) plus headers */
    return 4096 + (size_t)p->n_pics * ((size_t)p->width_mbs * p->height_mbs * 1800 + 256);
}

int mvs_generate(const mvs_params *p, mvs_output *out)
{
    if (!p || !out || p->width_mbs < 1 || p->height_mbs < 1 || p->n_pics < 1) return 0;
    if (p->profile_idc != 66 && p->profile_idc != 77 && p->profile_idc != 100) return 0;
    if (p->qp_min < 0 || p->qp_max > 51 || p->qp_min > p->qp_max) return 0;
    if (p->init_qp < p->qp_min || p->init_qp > p->qp_max) return 0;
    /* The reference parses the pic_order_cnt_type 1 fields for EVERY non-zero type
     * (h264_parameterset.c:314-331), so a conformant type-2 SPS desynchronises it. */
    if (p->poc_type != 0) return 0;

    enc_t e;
    memset(&e, 0, sizeof e);
    e.p = p; e.W = p->width_mbs; e.H = p->height_mbs;
    rng_seed(&e.rng, p->seed);
    make_zigzag(4, e.zz4); make_zigzag(8, e.zz8);
    for (int k = 0; k < 48; k++) e.cbp_to_codenum[cbp_intra_by_codenum[k]] = (uint8_t)k;

    size_t n_mb = (size_t)e.W * e.H;
    e.tot_luma = calloc(n_mb * 16, 1);
    e.tot_chroma[0] = calloc(n_mb * 4, 1);
    e.tot_chroma[1] = calloc(n_mb * 4, 1);
    e.mode_grid = calloc(n_mb * 16, 1);
    size_t pic_cap = n_mb * 1800 + 256;
    uint8_t *rbsp = malloc(pic_cap);
    uint8_t dummy[8];
    sink_t sink = { out->stream ? out->stream : dummy, out->stream ? out->stream_cap : 0, 0, 0 };
    int ok = e.tot_luma && e.tot_chroma[0] && e.tot_chroma[1] && e.mode_grid && rbsp;
    int high = p->profile_idc >= 100;

    /* scaling lists */
    for (int i = 0; i < 6; i++) memset(out->lists4x4[i], 16, 16);
    for (int i = 0; i < 2; i++) memset(out->lists8x8[i], 16, 64);
    if (high && p->scaling_lists) {
        for (int i = 0; i < 6; i++) for (int k = 0; k < 16; k++) out->lists4x4[i][k] = (uint8_t)(8 + rng_below(&e.rng, 41));
        for (int i = 0; i < 2; i++) for (int k = 0; k < 64; k++) out->lists8x8[i][k] = (uint8_t)(8 + rng_below(&e.rng, 41));
    }

    if (ok && out->stream) {
        /* ---- SPS (7.3.2.1.1) ---- */
        bw_t b = { rbsp, pic_cap, 0, 0, 0, 0 };
        bw_put(&b, 8, (uint32_t)p->profile_idc);
        bw_put(&b, 8, 0);                       /* constraint flags + reserved_zero_2bits */
        bw_put(&b, 8, 40);                      /* level_idc */
        bw_ue(&b, 0);                           /* seq_parameter_set_id */
        if (high) {
            bw_ue(&b, 1);                       /* chroma_format_idc 4:2:0 */
            bw_ue(&b, 0); bw_ue(&b, 0);         /* bit depths 8 */
            bw_put(&b, 1, 0);                   /* qpprime_y_zero_transform_bypass_flag */
            bw_put(&b, 1, p->scaling_lists ? 1 : 0);
            if (p->scaling_lists)
                for (int i = 0; i < 8; i++) {
                    bw_put(&b, 1, 1);           /* every list present (no fall-back in the reference) */
                    if (i < 6) put_scaling_list(&b, out->lists4x4[i], 16);
                    else       put_scaling_list(&b, out->lists8x8[i - 6], 64);
                }
        }
        bw_ue(&b, 0);                           /* log2_max_frame_num_minus4 */
        bw_ue(&b, (uint32_t)p->poc_type);
        if (p->poc_type == 0) bw_ue(&b, 0);     /* log2_max_pic_order_cnt_lsb_minus4 */
        bw_ue(&b, 1);                           /* max_num_ref_frames */
        bw_put(&b, 1, 0);                       /* gaps_in_frame_num_value_allowed_flag */
        bw_ue(&b, (uint32_t)(e.W - 1));
        bw_ue(&b, (uint32_t)(e.H - 1));
        bw_put(&b, 1, 1);                       /* frame_mbs_only_flag */
        bw_put(&b, 1, 1);                       /* direct_8x8_inference_flag */
        bw_put(&b, 1, p->crop_bottom ? 1 : 0);
        if (p->crop_bottom) { bw_ue(&b, 0); bw_ue(&b, 0); bw_ue(&b, 0); bw_ue(&b, (uint32_t)p->crop_bottom); }
        bw_put(&b, 1, 0);                       /* vui_parameters_present_flag */
        bw_trailing(&b);
        emit_nal(&sink, 0x67, &b);

        /* ---- PPS (7.3.2.2) ---- */
        bw_t c = { rbsp, pic_cap, 0, 0, 0, 0 };
        bw_ue(&c, 0); bw_ue(&c, 0);             /* pps id, sps id */
        bw_put(&c, 1, 0);                       /* entropy_coding_mode_flag: CAVLC */
        bw_put(&c, 1, 0);                       /* bottom_field_pic_order_in_frame_present_flag */
        bw_ue(&c, 0);                           /* num_slice_groups_minus1 */
        bw_ue(&c, 0); bw_ue(&c, 0);             /* num_ref_idx_l0/l1_default_active_minus1 */
        bw_put(&c, 1, 0); bw_put(&c, 2, 0);     /* weighted_pred_flag, weighted_bipred_idc */
        bw_se(&c, p->init_qp - 26);
        bw_se(&c, 0);                           /* pic_init_qs_minus26 */
        bw_se(&c, p->cb_qp_offset);
        bw_put(&c, 1, 1);                       /* deblocking_filter_control_present_flag */
        bw_put(&c, 1, 0);                       /* constrained_intra_pred_flag */
        bw_put(&c, 1, 0);                       /* redundant_pic_cnt_present_flag */
        if (high) {
            bw_put(&c, 1, p->transform8x8 ? 1 : 0);
            bw_put(&c, 1, 0);                   /* pic_scaling_matrix_present_flag */
            bw_se(&c, p->cr_qp_offset);
        }
        bw_trailing(&c);
        emit_nal(&sink, 0x68, &c);
        ok = !b.overflow && !c.overflow;
    }

    soa_t s = { out->mb_kind, out->i16_mode, out->chroma_mode, out->cbp, out->luma_modes, out->qp_y, out->coeff };
    for (int pic = 0; ok && pic < p->n_pics; pic++) {
        memset(e.tot_luma, 0, n_mb * 16);
        memset(e.tot_chroma[0], 0, n_mb * 4);
        memset(e.tot_chroma[1], 0, n_mb * 4);
        memset(e.mode_grid, 2, n_mb * 16);

        bw_t b = { rbsp, pic_cap, 0, 0, 0, 0 };
        /* ---- slice header (7.3.3), IDR I slice ---- */
        int span = p->qp_max - p->qp_min + 1;
        int slice_qp = p->qp_min + (int)rng_below(&e.rng, (uint32_t)span);
        bw_ue(&b, 0);                           /* first_mb_in_slice */
        bw_ue(&b, 7);                           /* slice_type: I (all slices) */
        bw_ue(&b, 0);                           /* pic_parameter_set_id */
        bw_put(&b, 4, 0);                       /* frame_num */
        bw_ue(&b, (uint32_t)(pic & 0xffff));    /* idr_pic_id */
        if (p->poc_type == 0) bw_put(&b, 4, 0); /* pic_order_cnt_lsb */
        bw_put(&b, 1, 0); bw_put(&b, 1, 0);     /* dec_ref_pic_marking: no_output.., long_term.. */
        bw_se(&b, slice_qp - p->init_qp);       /* slice_qp_delta */
        bw_ue(&b, 1);                           /* disable_deblocking_filter_idc */

        int qp_prev = slice_qp;
        for (int my = 0; my < e.H; my++)
            for (int mx = 0; mx < e.W; mx++)
                encode_mb(&e, &b, mx, my, &qp_prev, &s, (size_t)pic * n_mb + (size_t)my * e.W + mx);
        bw_trailing(&b);
        if (b.overflow) ok = 0;
        if (out->stream) emit_nal(&sink, 0x65, &b);
    }
    if (ok && out->stream) {
        for (int i = 0; i < 64; i++) sink_byte(&sink, 0);   /* esparser.c:65 stops 32 bytes early */
        if (sink.overflow) ok = 0;
        out->stream_len = sink.len;
    }
    free(e.tot_luma); free(e.tot_chroma[0]); free(e.tot_chroma[1]); free(e.mode_grid); free(rbsp);
    return ok;
}
