/*
 * recon_oracle.h -- TEST INFRASTRUCTURE.  CPU restatement (plain C) of the
 * reference's H.264 intra reconstruction + export path, operating on the
 * mvgpu.h structure-of-arrays.  It is the checker for the CUDA path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 * The product (libmvgpu.so) never links or calls it.
 *
 * PARITY PINNED: the reference ships no golden vectors, so this
 * restatement is pinned against outputs of the reference itself -- the
 * unmodified sources compiled into oracle/_ref (see oracle/Makefile) and run
 * on committed synthetic streams; byte-identical YUV and RGB on every fixture
 * in tests/golden/ (tests/test_oracle_vs_reference.py regenerates the check
 * whenever the reference tree is mounted).
 */
#ifndef RECON_ORACLE_H
#define RECON_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_sps {
    int32_t width_mbs, height_mbs;
    int32_t level_scale4x4[3][6][16];   /* [Y,Cb,Cr][qP%6][i*4+j] */
    int32_t level_scale8x8[6][64];      /* [qP%6][i*8+j]          */
    int32_t cb_qp_offset, cr_qp_offset;
} oracle_sps;

/* h264.c:419-493 (normAdjust) + h264_parameterset.c:280-303 +
 * h264_transform.c:645-741 (LevelScale).  lists are zig-zag ordered scaling
 * lists, NULL = flat 16. */
void oracle_build_level_scale(const uint8_t *lists4x4 /*[3][16]*/, const uint8_t *list8x8 /*[64]*/,
                              int32_t ls4[3][6][16], int32_t ls8[6][64]);

/* h264_transform.c:598-637 */
int oracle_chroma_qp(int qp_y, int offset);

/* Dequantisation + inverse transforms of one macroblock (kernel 1).
 * residual: [0..255] luma raster (y*16+x), [256..319] Cb (y*8+x), [320..383] Cr.
 * h264_transform.c:121-402, :756-860, :1100-1383. */
void oracle_mb_residual(const oracle_sps *sps, int mb_kind, int qp_y,
                        const int16_t coeff[384], int32_t residual[384]);

/* Reconstruct one picture (kernels 1+2).  Arrays are indexed by mbAddr as in
 * mvgpu.h.  y: W*H, cb/cr: (W/2)*(H/2).  residual_out (optional): [N][384]
 * int16, clamped to [-512, 511] (the intermediate kernel 1 hands to kernel 2). */
void oracle_reconstruct_picture(const oracle_sps *sps,
                                const uint8_t *mb_kind, const uint8_t *i16_mode,
                                const uint8_t *chroma_mode, const int8_t *qp_y,
                                const uint8_t *luma_modes, const int16_t *coeff,
                                uint8_t *y, uint8_t *cb, uint8_t *cr,
                                int16_t *residual_out);

/* export_utils.c:209-324 (mb_to_rgb) for scale == 1; for scale s > 1 the
 * rounded s x s box average of that RGB picture (no reference counterpart,
 * SURVEY.md section 8 row a32). rgb: (W/s)*(H/s)*3. */
void oracle_yuv420_to_rgb(int width, int height, const uint8_t *y, const uint8_t *cb,
                          const uint8_t *cr, int scale, uint8_t *rgb);

#ifdef __cplusplus
}
#endif
#endif
