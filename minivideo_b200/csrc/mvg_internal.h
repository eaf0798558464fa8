/*
 * mvg_internal.h -- layouts shared by the CUDA kernels (mvg_kernels.cuh) and the
 * host side of libmvgpu.so (mvg_api.cu).  Not part of the public ABI.
 */
#ifndef MVG_INTERNAL_H
#define MVG_INTERNAL_H

#include <stdint.h>

/* ---- kernel 2 shared-memory tile geometry (per warp) ----------------------
 * luma tile  : rows y = -1..15, row stride 32 B, sample x at byte (x + 8):
 *              x = -1 -> 7, x = 0..15 -> 8..23, x = 16..23 (MB C) -> 24..31
 * chroma tile: rows y = -1..7, row stride 16 B, sample x at byte (x + 8)        */
#define MVG_LT_STRIDE 32
#define MVG_LT_XOFF   8
#define MVG_LT_ROWS   17
#define MVG_CT_STRIDE 16
#define MVG_CT_XOFF   8
#define MVG_CT_ROWS   9

/* ---- per-macroblock control record kernel 1 writes for kernel 2 (16 B) ----
 * w0: byte0 mb_kind, byte1 Intra16x16PredMode, byte2 intra_chroma_pred_mode,
 *     byte3 reserved
 * w1,w2: 16 luma prediction modes, 4 bits each (block b at bits 4b of the
 *     64-bit value w1 | w2<<32); Intra8x8 modes sit in nibbles 0..3
 * w3: bit b set <=> 4x4 block b has a non-zero residual sample
 *     (b = 0..15 luma blocks in decoding order, 16..19 Cb, 20..23 Cr)          */
struct MvgMbCtl { uint32_t w0, w1, w2, w3; };

/* ---- prediction tap tables ---------------------------------------------------
 * Every directional Intra4x4 / Intra8x8 predictor is
 *     pred = (n[a] + n[b] + n[c] + n[d] + 2) >> 2
 * over four (possibly repeated) neighbour samples:
 *   3-tap (p + 2q + r + 2) >> 2  -> {p,q,q,r}
 *   2-tap (p + q + 1) >> 1       -> {p,p,q,q}
 *   copy  p                      -> {p,p,p,p}
 *   end   (p + 3q + 2) >> 2      -> {p,q,q,q}
 * lut4[tr][mode][y*4+x]: four uint8 byte offsets into the luma tile relative to
 *   (the block's top-left sample - MVG_LUT4_BIAS); tr = 1 when p[4..7,-1] are
 *   available, else they alias p[3,-1] (h264_intra_prediction.c:431-439).
 *   Mode 2 (DC) is unused.
 * lut8[mode][y*8+x]: the Intra8x8 predictors read a 26-entry filtered neighbour line
 *   (0..7 = p'[-1,7..0], 8 = p'[-1,-1], 9..24 = p'[0..15,-1], 25 = the DC value), each
 *   entry a 32-bit word {p', f2 = (p'[n]+p'[n+1]+1)>>1, f3 = (p'[n-1]+2p'[n]+p'[n+1]+2)>>2};
 *   a sample is one byte of one word: low byte = 4*n (byte offset of the word), high byte =
 *   bit shift (0 copy, 8 two-tap, 16 three-tap).  The 3- and 2-tap forms of the spec always
 *   involve adjacent line entries, and the two "end" taps (p+3q) are f3 at a line end.      */
#define MVG_LUT4_BIAS    (MVG_LT_STRIDE + 1)   /* p[-1,-1] is the lowest address */
#define MVG_N8_LEFT(y)  (7 - (y))
#define MVG_N8_CORNER   8
#define MVG_N8_TOP(x)   (9 + (x))
#define MVG_N8_DC       25

struct MvgLuts {
    uint32_t lut4[2][9][16];
    uint16_t lut8[9][64];
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
