/*
 * ref_driver.c -- TEST INFRASTRUCTURE.  Drives the UNMODIFIED reference decoder
 * (objects compiled by oracle/Makefile from /root/reference/minivideo/src where
 * they lie) through its public API and taps every decoded IDR picture.
 *
 * The reference exports pictures only as files in the CWD (export.c:618-767).
 * This driver is linked with `-Wl,--wrap=export_idr`, so decode_slice()'s call
 * (h264_slice.c:96-99) lands in __wrap_export_idr() below while every reference
 * source stays untouched.  For each picture it can write
 *   --yuv  F : planar I420 gathered by the reference's own mb_to_ycbcr()
 *              (export_utils.c:117) -- same bytes as export_idr_yuv420()
 *   --rgb  F : RGB24 produced by the reference's own mb_to_rgb() (export_utils.c:209)
 *   --soa  F : the reference's parsed Macroblock_t records converted to the
 *              mvgpu.h structure-of-arrays (golden INPUT of the GPU boundary)
 *   --export : additionally call the real export_idr() (files in the CWD)
 *   --time   : print wall seconds spent inside minivideo_decode(); with --time
 *              and no file outputs the tap still runs mb_to_rgb() into a scratch
 *              buffer so the timed work is parse + reconstruction + RGB.
 *
 * usage: ref_decode <in.264> <n_pictures> [--yuv F] [--rgb F] [--soa F] [--export] [--time] [--norgb]
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "minivideo.h"
#include "export.h"
#include "export_utils.h"
#include "utils.h"
#include "decoder/h264/h264_decodingcontext.h"

int __real_export_idr(DecodingContext_t *dc);

static FILE *f_yuv, *f_rgb, *f_soa;
static int want_export, want_time, want_rgb = 1;
static int n_tapped;
static unsigned char *scratch;
static size_t scratch_size;

/* File layout of --soa (little endian):
 *   int32 magic 'MVSA', w_mbs, h_mbs, n_pics (patched at exit), cb_off, cr_off,
 *   int32 ls4[3][6][16], int32 ls8[6][64]
 *   then per picture: mb_kind[N] i16_mode[N] chroma_mode[N] qp_y[N] cbp[N]
 *                     luma_modes[N*16] coeff int16[N*384]                    */
static void soa_header(DecodingContext_t *dc)
{
    pps_t *pps = dc->pps_array[dc->active_slice->pic_parameter_set_id];
    sps_t *sps = dc->sps_array[pps->seq_parameter_set_id];
    int32_t h[6] = {0x4153564d, (int32_t)sps->PicWidthInMbs, (int32_t)sps->PicHeightInMapUnits, 0,
                    pps->chroma_qp_index_offset, pps->second_chroma_qp_index_offset};
    fwrite(h, 4, 6, f_soa);
    for (int c = 0; c < 3; c++)
        for (int q = 0; q < 6; q++)
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 4; j++) {
                    int32_t v = sps->LevelScale4x4[c][q][i][j];
                    fwrite(&v, 4, 1, f_soa);
                }
    for (int q = 0; q < 6; q++)
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++) {
                int32_t v = sps->LevelScale8x8[0][q][i][j];
                fwrite(&v, 4, 1, f_soa);
            }
}

static void soa_picture(DecodingContext_t *dc)
{
    unsigned n = dc->PicSizeInMbs;
    uint8_t *kind = calloc(n, 1), *i16 = calloc(n, 1), *cm = calloc(n, 1), *cbp = calloc(n, 1);
    int8_t *qp = calloc(n, 1);
    uint8_t *modes = calloc(n, 16);
    int16_t *coeff = calloc((size_t)n * 384, 2);

    for (unsigned a = 0; a < n; a++) {
        Macroblock_t *mb = dc->mb_array[a];
        if (!mb) continue;
        int16_t *c = coeff + (size_t)a * 384;
        unsigned m = mb->MbPartPredMode[0];
        qp[a] = (int8_t)mb->QPY;
        cm[a] = (uint8_t)mb->IntraChromaPredMode;
        cbp[a] = (uint8_t)((mb->CodedBlockPatternChroma << 4) | mb->CodedBlockPatternLuma);
        if (m == Intra_4x4) {
            kind[a] = 0;
            for (int b = 0; b < 16; b++) {
                modes[a * 16 + b] = (uint8_t)mb->Intra4x4PredMode[b];
                for (int k = 0; k < 16; k++) c[b * 16 + k] = (int16_t)mb->LumaLevel4x4[b][k];
            }
        } else if (m == Intra_8x8) {
            kind[a] = 1;
            for (int b = 0; b < 4; b++) {
                modes[a * 16 + b] = (uint8_t)mb->Intra8x8PredMode[b];
                for (int k = 0; k < 64; k++) c[b * 64 + k] = (int16_t)mb->LumaLevel8x8[b][k];
            }
        } else {
            kind[a] = 2;
            i16[a] = (uint8_t)mb->Intra16x16PredMode;
            int dcm[4][4];
            inverse_scan_4x4(mb->Intra16x16DCLevel, dcm);   /* h264_transform.c:180 */
            for (int b = 0; b < 16; b++) {
                c[b * 16] = (int16_t)dcm[raster_4x4_2d[b][0]][raster_4x4_2d[b][1]];
                for (int k = 1; k < 16; k++) c[b * 16 + k] = (int16_t)mb->Intra16x16ACLevel[b][k - 1];
            }
        }
        for (int p = 0; p < 2; p++)
            for (int b = 0; b < 4; b++) {
                c[256 + p * 64 + b * 16] = (int16_t)mb->ChromaDCLevel[p][b];
                for (int k = 1; k < 16; k++)
                    c[256 + p * 64 + b * 16 + k] = (int16_t)mb->ChromaACLevel[p][b][k - 1];
            }
    }
    fwrite(kind, 1, n, f_soa); fwrite(i16, 1, n, f_soa); fwrite(cm, 1, n, f_soa);
    fwrite(qp, 1, n, f_soa);   fwrite(cbp, 1, n, f_soa); fwrite(modes, 16, n, f_soa);
    fwrite(coeff, 2, (size_t)n * 384, f_soa);
    free(kind); free(i16); free(cm); free(cbp); free(qp); free(modes); free(coeff);
}

int __wrap_export_idr(DecodingContext_t *dc)
{
    sps_t *sps = dc->sps_array[dc->active_sps];
    size_t w = sps->PicWidthInMbs * 16, h = sps->PicHeightInMapUnits * 16;
    if (scratch_size < w * h * 3) {
        free(scratch);
        scratch_size = w * h * 3;
        scratch = malloc(scratch_size);
    }
    if (f_soa) {
        if (n_tapped == 0) soa_header(dc);
        soa_picture(dc);
    }
    if (f_yuv) {
        mb_to_ycbcr(dc, scratch);
        fwrite(scratch, 1, w * h * 3 / 2, f_yuv);
    }
    if (f_rgb || (want_time && want_rgb)) {
        mb_to_rgb(dc, scratch);
        if (f_rgb) fwrite(scratch, 1, w * h * 3, f_rgb);
    }
    n_tapped++;
    if (want_export) return __real_export_idr(dc);
    dc->picture_exported++;
    return SUCCESS;
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s <in.264> <n_pictures> [--yuv F] [--rgb F] [--soa F] [--export] [--time] [--norgb]\n", argv[0]);
        return 2;
    }
    int n = atoi(argv[2]);
    for (int i = 3; i < argc; i++) {
        if (!strcmp(argv[i], "--yuv") && i + 1 < argc) f_yuv = fopen(argv[++i], "wb");
        else if (!strcmp(argv[i], "--rgb") && i + 1 < argc) f_rgb = fopen(argv[++i], "wb");
        else if (!strcmp(argv[i], "--soa") && i + 1 < argc) f_soa = fopen(argv[++i], "wb");
        else if (!strcmp(argv[i], "--export")) want_export = 1;
        else if (!strcmp(argv[i], "--time")) want_time = 1;
        else if (!strcmp(argv[i], "--norgb")) want_rgb = 0;
    }

    MediaFile_t *media = NULL;
    int rc = minivideo_open(argv[1], &media);
    if (rc != SUCCESS) { fprintf(stderr, "ref_decode: open failed\n"); return 1; }
    rc = minivideo_parse(media, false, true, false);
    if (rc != SUCCESS) { fprintf(stderr, "ref_decode: parse failed\n"); return 1; }

    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    rc = minivideo_decode(media, ".", PICTURE_YUV420, 75, n, PICTURE_UNFILTERED);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    minivideo_close(&media);

    if (f_soa) {
        int32_t np = n_tapped;
        fseek(f_soa, 12, SEEK_SET);
        fwrite(&np, 4, 1, f_soa);
        fclose(f_soa);
    }
    if (f_yuv) fclose(f_yuv);
    if (f_rgb) fclose(f_rgb);
    if (want_time)
        printf("REFTIME pictures=%d seconds=%.6f\n", n_tapped,
               (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec));
    else
        printf("REFDONE pictures=%d rc=%d\n", n_tapped, rc);
    return n_tapped == n ? 0 : 3;
}
