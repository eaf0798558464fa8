/*
 * mvsynth.h -- synthetic intra-only H.264 stream generator (libmvsynth.so).
 *
 * A minimal CAVLC intra ENCODER with random modes and residuals.  It does no
 * rate/distortion work and never reconstructs: it draws syntax elements from a
 * seeded RNG and writes them as a legal Annex-B byte stream, and at the same
 * time emits the very same syntax as the mvgpu.h structure-of-arrays.  The
 * stream obeys every parsing quirk of the reference decoder listed in
 * SURVEY.md section 8(c) (4-byte start codes, SPS/PPS ids 0, one slice per picture,
 * all 8 SPS scaling lists present, no QP'Y == 36 on Intra16x16 MBs, only
 * prediction modes whose neighbours exist, ...), so the reference decodes what
 * the SoA says.
 *
 * It is benchmark/test input tooling: nothing in libmvgpu.so depends on it.
 */
#ifndef MVSYNTH_H
#define MVSYNTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mvs_params {
    int32_t  width_mbs, height_mbs;
    int32_t  n_pics;
    int32_t  profile_idc;       /* 66 Baseline, 77 Main, 100 High                       */
    int32_t  transform8x8;      /* High only: transform_8x8_mode_flag, allows I8x8 MBs  */
    int32_t  scaling_lists;     /* High only: 8 random SPS scaling lists in [8,48]      */
    int32_t  w_i4x4, w_i8x8, w_i16x16;  /* relative MB type weights                     */
    int32_t  init_qp;           /* pic_init_qp (26 + pic_init_qp_minus26)               */
    int32_t  qp_min, qp_max;    /* QPY is kept inside [qp_min, qp_max]                  */
    int32_t  cb_qp_offset, cr_qp_offset; /* cr only signalled for High                  */
    int32_t  luma_cbp_percent;  /* P(8x8 luma cbp bit) in percent                        */
    int32_t  mean_coeffs_x10;   /* mean TotalCoeff of a coded 4x4 block, times 10        */
    int32_t  level_scale_x10;   /* Laplacian scale of |level|-1, times 10                */
    int32_t  max_level;         /* |level| <= max_level (AC); DC limited to 4*max_level  */
    int32_t  poc_type;          /* must be 0 (the reference mis-parses types 1 and 2)     */
    int32_t  crop_bottom;       /* frame_crop_bottom_offset in luma rows/2 (signalled only) */
    int32_t  force_mode;        /* -1 random; else use this pred mode wherever legal      */
    int32_t  force_kind;        /* -1 random; else 0/1/2                                  */
    uint64_t seed;
} mvs_params;

/* Output of one generation call.  SoA arrays follow mvgpu.h (mvg_batch) and are
 * sized for n_pics * width_mbs * height_mbs macroblocks; the caller allocates.
 * Any SoA pointer may be NULL (then only the stream is produced); stream may be
 * NULL (then only the SoA is produced). */
typedef struct mvs_output {
    uint8_t *stream;  size_t stream_cap;  size_t stream_len;
    uint8_t *mb_kind, *i16_mode, *chroma_mode, *cbp, *luma_modes;
    int8_t  *qp_y;
    int16_t *coeff;
    uint8_t  lists4x4[6][16];   /* zig-zag scaling lists actually signalled (16 = flat) */
    uint8_t  lists8x8[2][64];
} mvs_output;

void mvs_default_params(mvs_params *p);

/* Returns 1 on success, 0 on failure (buffer too small, bad params). */
int mvs_generate(const mvs_params *p, mvs_output *out);

/* Upper bound of the stream size for these params (bytes). */
size_t mvs_stream_bound(const mvs_params *p);

#ifdef __cplusplus
}
#endif
#endif
