/*
 * mvg_internal.h -- layouts shared by the CUDA kernels (mvg_kernels.cuh) and the
 * host side of libmvgpu.so (mvg_api.cu).  Not part of the public ABI.
 */
#ifndef MVG_INTERNAL_H
#define MVG_INTERNAL_H

#include <stdint.h>

/* ---- kernel 2 shared-memory tile geometry (per warp) ----------------------
 * luma tile  : rows y = -1..15, sample x at byte (x + 16): x = -1 -> 15, x = 0..15 -> 16..31,
 *              x = 16..23 (macroblock C, top row only) -> 32..39.  Row stride 40 B = 10 words: 10 y mod 32 is
 *              a different even bank for each of 16 rows, so a column (x = -1 hand-over, Horizontal
 *              predictors) and the 8-byte row pieces of the write-out are free of bank conflicts; rows are
 *              8-byte aligned.
 * chroma tile: rows y = -1..7, sample x at byte (x + 8), row stride 24 B = 6 words, Cr plane 320 B = 80 words
 *              after Cb: the 2 x 8 rows x 2 words of both planes cover all 32 banks exactly once. */
#ifndef MVG_LT_STRIDE
#define MVG_LT_STRIDE 40
#endif
#define MVG_LT_XOFF   16
#define MVG_LT_ROWS   17
#define MVG_CT_STRIDE 24
#define MVG_CT_XOFF   8
#define MVG_CT_ROWS   9
#define MVG_CT_PLANE  320

/* ---- per-macroblock control record kernel 1 writes for kernel 2 (16 B) ----
 * w0: byte0 mb_kind, byte1 Intra16x16PredMode, byte2 intra_chroma_pred_mode,
 *     byte3 reserved
 * w1,w2: 16 luma prediction modes, 4 bits each, in the order kernel 2's anti-diagonal Intra4x4
 *     schedule consumes them (step t predicts block (t&1, t>>1) in lanes 0..15 and block
 *     ((t&1)+2, (t>>1)-1) in lanes 16..31):
 *       w1 nibble t (t = 0..7)  = mode of 4x4 block (t&1, t>>1)        = blkIdx 0,1,2,3,8,9,10,11
 *       w2 = rotl(v, 8), v nibble t-2 (t = 2..9) = mode of block ((t&1)+2, (t>>1)-1)
 *                                                                   = blkIdx 4,5,6,7,12,13,14,15
 *     so that both lane halves find the mode of step t at nibble (t & 7).
 *     Intra8x8 modes sit in nibbles 0..3 of w1
 * w3: reserved (0)                                                             */
struct MvgMbCtl { uint32_t w0, w1, w2, w3; };

/* ---- prediction tables -----------------------------------------------------
 * Every directional Intra4x4 / Intra8x8 predictor is
 *     pred = (n[a] + n[b] + n[c] + n[d] + 2) >> 2
 * over four (possibly repeated) neighbour samples:
 *   3-tap (p + 2q + r + 2) >> 2  -> {p,q,q,r}
 *   2-tap (p + q + 1) >> 1       -> {p,p,q,q}
 *   copy  p                      -> {p,p,p,p}
 *   end   (p + 3q + 2) >> 2      -> {p,q,q,q}
 * lut4[mode][16*half + y*4+x]: the four taps as byte offsets into the luma tile (one byte each), relative
 *   to (origin of the lane's block) - MVG_LUT4_BIAS; the two halves hold the same values (a lane reads
 *   word `lane`: no bank conflicts).  Rows 0..8 are the nine modes with p[4..7,-1] available, rows 11 and
 *   15 are modes 3 and 7 when they are not (taps stop at p[3,-1], h264_intra_prediction.c:431-439).
 *   DC: row 9 = the four samples above, row 10 = the four to the left (one side available: (sum + 2) >> 2 like
 *   any other row); row 2 (both sides) = lanes with even x the four above, odd x the four to the left, the
 *   kernel adds the half-sum of lane ^ 1 and shifts by 3.
 * lut8[mode][lane]: the Intra8x8 predictors read a filtered neighbour line kept as 32 words of three bytes, one word per
 *   line entry n (n = 0..7 p'[-1,7..0], 8 p'[-1,-1], 9..24 p'[0..15,-1]), written with ONE store per lane:
 *   byte 0 p', byte 1 f2 = (p'[n]+p'[n+1]+1)>>1, byte 2 f3 = (p'[n-1]+2p'[n]+p'[n+1]+2)>>2; byte MVG_N8_DC, behind
 *   the 32 words, holds the DC value.  A lane predicts two horizontally adjacent
 *   samples, x = 4*((lane>>3)&1) + 2*(lane&1) + {0,1}, y = 4*(lane>>4) + ((lane>>1)&3) (residual word 4*lane of the
 *   block): bits 0..15 = byte index (4 * n + variant) of sample 0, bits 16..31 = of sample 1.  The 3- and 2-tap forms
 *   of the spec always involve adjacent line entries, and the two "end" taps (p+3q) are f3 at a line end. */
#define MVG_LUT4_BIAS    (MVG_LT_STRIDE + 1)   /* p[-1,-1] is the lowest address */
#define MVG_N8_LEFT(y)  (7 - (y))
#define MVG_N8_CORNER   8
#define MVG_N8_TOP(x)   (9 + (x))
#define MVG_N8_IDX(n, variant) (4 * (n) + (variant))
#define MVG_N8_DC       128     /* byte index of the DC value behind the 32 line words */
#define MVG_N8_BYTES    144

struct MvgLuts {
    uint32_t lut4[16][32];
    uint32_t lut8[16][32];      /* rows 9..15 are zero: a mode nibble outside 0..8 (malformed input) reads entry 0, never out of bounds */
};

#ifdef __cplusplus
extern "C" {
#endif
/* mvg_tables.cpp-style host helpers implemented in mvg_api.cu */
void mvg_build_luts(struct MvgLuts *out);
#ifdef __cplusplus
}
#endif

#endif
