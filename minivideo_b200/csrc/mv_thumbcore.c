/*
 * mv_thumbcore.c -- thumbnail extraction over the GPU path, shared by the mv_thumbnailer CLI and by the
 * drop-in public API (minivideo_shim.c):
 *
 * Annex-B bytes -> mvf_open_annexb() -> mvf_select_idr() (demuxer/filter.c semantics) ->
 * mvf_parse_pictures_packed() (threaded CAVLC, packed levels) -> mvg_decode_host_packed() (pinned copies +
 * kernels 0-4) -> picture
 * files named like export_idr() names them (export.c:627-642,:704-705): <input base name>[_<k>].<ext>
 * with k counting exported pictures when more than one was requested.  File contents are byte-identical
 * to the reference's: planar I420 (export.c:100-151), 24-bit bottom-up BMP and run-length TGA as
 * stb_image_write lays them out (export.c:535-539,:566-570).
 * No CPU fallback: without a CUDA device mvt_extract() fails.
 */
#include <limits.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "mvfront.h"
#include "mvgpu.h"
#include "mv_thumbcore.h"


static void put16(FILE *f, unsigned v) { fputc(v & 255, f); fputc((v >> 8) & 255, f); }
static void put32(FILE *f, unsigned v) { put16(f, v & 0xffff); put16(f, v >> 16); }

/* 24-bit uncompressed BMP: 14-byte file header, 40-byte BITMAPINFOHEADER, rows bottom-up, BGR,
 * padded to 4 bytes -- the layout stbi_write_bmp() produces for 3 components */
static int write_bmp(const char *path, const uint8_t *rgb, int w, int h)
{
    FILE *f = fopen(path, "wb");
    if (!f) return 0;
    int pad = (4 - (w * 3) % 4) % 4;
    fputc('B', f); fputc('M', f);
    put32(f, (unsigned)(54 + (w * 3 + pad) * h)); put16(f, 0); put16(f, 0); put32(f, 54);
    put32(f, 40); put32(f, (unsigned)w); put32(f, (unsigned)h); put16(f, 1); put16(f, 24);
    for (int i = 0; i < 6; i++) put32(f, 0);
    uint8_t *line = malloc((size_t)w * 3 + 4);
    for (int y = h - 1; y >= 0; y--) {
        const uint8_t *src = rgb + (size_t)y * w * 3;
        for (int x = 0; x < w; x++) { line[3 * x] = src[3 * x + 2]; line[3 * x + 1] = src[3 * x + 1]; line[3 * x + 2] = src[3 * x]; }
        memset(line + w * 3, 0, (size_t)pad);
        fwrite(line, 1, (size_t)(w * 3 + pad), f);
    }
    free(line);
    return fclose(f) == 0;
}

static int same_px(const uint8_t *a, const uint8_t *b) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2]; }
static void put_bgr(FILE *f, const uint8_t *p) { fputc(p[2], f); fputc(p[1], f); fputc(p[0], f); }

/* run-length true-colour TGA (image type 10), origin bottom-left, BGR.  Packets are formed the way
 * stbi_write_tga() forms them so the files compare equal: a raw packet keeps growing while pixel k
 * differs from pixel k-2 and gives its last pixel back when they match; a run packet grows while
 * pixels equal its first one; both stop at 128 pixels. */
static int write_tga(const char *path, const uint8_t *rgb, int w, int h)
{
    FILE *f = fopen(path, "wb");
    if (!f) return 0;
    fputc(0, f); fputc(0, f); fputc(10, f);
    put16(f, 0); put16(f, 0); fputc(0, f);
    put16(f, 0); put16(f, 0); put16(f, (unsigned)w); put16(f, (unsigned)h);
    fputc(24, f); fputc(0, f);
    for (int y = h - 1; y >= 0; y--) {
        const uint8_t *row = rgb + (size_t)y * w * 3;
        int len;
        for (int i = 0; i < w; i += len) {
            const uint8_t *first = row + 3 * i;
            int is_run = 0;
            len = 1;
            if (i < w - 1) {
                len = 2;
                is_run = same_px(first, first + 3);
                if (is_run) {
                    for (int k = i + 2; k < w && len < 128 && same_px(first, row + 3 * k); k++) len++;
                } else {
                    for (int k = i + 2; k < w && len < 128; k++) {
                        if (same_px(row + 3 * (k - 2), row + 3 * k)) { len--; break; }
                        len++;
                    }
                }
            }
            if (is_run) { fputc(len + 127, f); put_bgr(f, first); }
            else { fputc(len - 1, f); for (int k = 0; k < len; k++) put_bgr(f, first + 3 * k); }
        }
    }
    return fclose(f) == 0;
}

static int write_raw(const char *path, const uint8_t *data, size_t n)
{
    FILE *f = fopen(path, "wb");
    if (!f) return 0;
    size_t w = fwrite(data, 1, n, f);
    return fclose(f) == 0 && w == n;
}

/* one feeder per GPU: its own context, pinned buffers and parser threads; it takes every
 * n_gpus-th batch of the selected pictures, no data crosses GPUs */
typedef struct {
    mvf_stream *st; const mvf_info *info; const int32_t *sel; int n_sel;
    const char *base, *outdir; int fmt, scale, device, threads, batch, numbered;
    int first_batch, batch_stride;      /* batches first_batch, first_batch + batch_stride, ... */
    int exported, rc;
} feeder_t;

static void *feeder_main(void *arg)
{
    feeder_t *f = arg;
    const mvf_info *info = f->info;
    const int W = 16 * info->width_mbs, H = 16 * info->height_mbs, batch = f->batch, fmt = f->fmt, scale = f->scale;
    static const char *ext[] = {"yuv", "bmp", "tga"};
    f->rc = 1;
    mvg_ctx *ctx = NULL;
    if (mvg_create(&ctx, f->device, info->width_mbs, info->height_mbs, batch) != MVG_SUCCESS) {
        fprintf(stderr, "mvt_extract: %s\n", mvg_last_error(NULL));
        return NULL;
    }
    if (mvg_set_sps(ctx, info->width_mbs, info->height_mbs, info->level_scale4x4, info->level_scale8x8,
                    info->cb_qp_offset, info->cr_qp_offset) != MVG_SUCCESS) {
        fprintf(stderr, "mvt_extract: %s\n", mvg_last_error(ctx));
        mvg_destroy(ctx);
        return NULL;
    }
    const size_t N = (size_t)info->width_mbs * info->height_mbs, nb = N * (size_t)batch;
    /* parsed pictures travel in the packed transfer format (mvgpu.h): a fifth of the dense levels on the bus */
    mvf_packed_batch pb;
    pb.n_pics = 0;
    pb.mb_kind = mvg_host_alloc(nb); pb.i16_mode = mvg_host_alloc(nb); pb.chroma_mode = mvg_host_alloc(nb);
    pb.qp_y = mvg_host_alloc(nb); pb.luma_modes = mvg_host_alloc(nb * 16);
    pb.nz_blocks = mvg_host_alloc(nb * sizeof(uint32_t)); pb.word_off = mvg_host_alloc(nb * sizeof(uint32_t));
    pb.pic_off = mvg_host_alloc(((size_t)batch + 1) * sizeof(uint64_t));
    pb.words_capacity = nb * MVG_PACKED_WORDS_PER_MB;
    pb.words = mvg_host_alloc(pb.words_capacity * sizeof(uint16_t));
    const int ow = W / scale, oh = H / scale;
    const size_t yuv_sz = (size_t)W * H * 3 / 2, rgb_sz = (size_t)ow * oh * 3;
    uint8_t *out = mvg_host_alloc((fmt == MVT_YUV420 ? yuv_sz : rgb_sz) * (size_t)batch);
    int rc = 0;
    if (!pb.mb_kind || !pb.i16_mode || !pb.chroma_mode || !pb.qp_y || !pb.luma_modes || !pb.nz_blocks || !pb.word_off ||
        !pb.pic_off || !pb.words || !out) {
        fprintf(stderr, "mvt_extract: pinned host allocation failed\n");
        rc = 1;
    }
    for (int bi = f->first_batch; !rc && bi * batch < f->n_sel; bi += f->batch_stride) {
        const int done = bi * batch;
        int cnt = f->n_sel - done < batch ? f->n_sel - done : batch;
        if (mvf_parse_pictures_packed(f->st, f->sel + done, 0, cnt, &pb, f->threads) != MVG_SUCCESS) {
            fprintf(stderr, "mvt_extract: %s\n", mvf_last_error(f->st));
            rc = 1; break;
        }
        mvg_packed_batch gb = { cnt, pb.mb_kind, pb.i16_mode, pb.chroma_mode, pb.qp_y, pb.luma_modes,
                                pb.nz_blocks, pb.word_off, pb.pic_off, pb.words };
        int ok = fmt == MVT_YUV420 ? mvg_decode_host_packed(ctx, &gb, out, NULL, 0) : mvg_decode_host_packed(ctx, &gb, NULL, out, scale);
        if (ok != MVG_SUCCESS) { fprintf(stderr, "mvt_extract: %s\n", mvg_last_error(ctx)); rc = 1; break; }
        for (int k = 0; k < cnt && !rc; k++) {
            char path[PATH_MAX];
            /* export_idr() numbers pictures in export order (export.c:630): the position in the selection */
            if (f->numbered) snprintf(path, sizeof path, "%s/%s_%d.%s", f->outdir, f->base, done + k, ext[fmt]);
            else snprintf(path, sizeof path, "%s/%s.%s", f->outdir, f->base, ext[fmt]);
            int w = fmt == MVT_YUV420 ? write_raw(path, out + (size_t)k * yuv_sz, yuv_sz)
                  : fmt == MVT_BMP    ? write_bmp(path, out + (size_t)k * rgb_sz, ow, oh)
                                      : write_tga(path, out + (size_t)k * rgb_sz, ow, oh);
            if (!w) { fprintf(stderr, "mvt_extract: cannot write '%s'\n", path); rc = 1; }
            else f->exported++;
        }
    }
    mvg_host_free(pb.mb_kind); mvg_host_free(pb.i16_mode); mvg_host_free(pb.chroma_mode); mvg_host_free(pb.qp_y);
    mvg_host_free(pb.luma_modes); mvg_host_free(pb.nz_blocks); mvg_host_free(pb.word_off); mvg_host_free(pb.pic_off);
    mvg_host_free(pb.words); mvg_host_free(out);
    mvg_destroy(ctx);
    f->rc = rc;
    return NULL;
}

int mvt_extract(const uint8_t *data, size_t len, const char *base, const char *outdir, int fmt, int n_want, int mode,
                int scale, int device, int threads, int batch, int *n_exported)
{
    if (n_exported) *n_exported = 0;
    if (!data || !base || n_want < 1 || scale < 1 || batch < 1 || fmt < MVT_YUV420 || fmt > MVT_TGA) return MVG_FAILURE;
    if (!outdir || !*outdir) outdir = ".";
    if (threads < 1) { long c = sysconf(_SC_NPROCESSORS_ONLN); threads = c > 0 ? (int)c : 1; }
    mvf_stream *st = NULL;
    if (mvf_open_annexb(data, len, &st) != MVG_SUCCESS) {
        fprintf(stderr, "mvt_extract: %s\n", mvf_last_error(NULL));
        return MVG_FAILURE;
    }
    mvf_info info;
    mvf_get_info(st, &info);
    int32_t *sel = malloc(sizeof(int32_t) * (size_t)(n_want > info.n_idr ? n_want : info.n_idr + 1));
    int n_sel = mvf_select_idr(st, n_want, mode, sel);
    if (n_sel < 1) { fprintf(stderr, "mvt_extract: no picture to decode after filtering\n"); free(sel); mvf_close(st); return MVG_FAILURE; }
    /* the reference appends _<k> when more than one picture was requested after filtering (export.c:630) */
    int numbered;
    if (mode == 0) numbered = (n_want < info.n_idr ? n_want : info.n_idr) > 1;
    else {      /* picture_number after filtering = min(requested, candidates) = what 'ordered' would return */
        int32_t *tmp = malloc(sizeof(int32_t) * (size_t)(n_want > info.n_idr ? n_want : info.n_idr + 1));
        numbered = mvf_select_idr(st, n_want, 1, tmp) > 1;
        free(tmp);
    }
    const int W = 16 * info.width_mbs, H = 16 * info.height_mbs;
    if (fmt != MVT_YUV420 && (W % scale || H % scale)) {
        fprintf(stderr, "mvt_extract: scale %d does not divide %dx%d\n", scale, W, H);
        free(sel); mvf_close(st);
        return MVG_FAILURE;
    }
    if (batch > n_sel) batch = n_sel;

    /* device >= 0: that GPU; device < 0: all of them, batches dealt round-robin */
    int n_gpus = 1;
    if (device < 0) {
        n_gpus = mvg_device_count();
        if (n_gpus < 1) { fprintf(stderr, "mvt_extract: no CUDA device; this path has no CPU fallback\n"); free(sel); mvf_close(st); return MVG_FAILURE; }
        const int n_batches = (n_sel + batch - 1) / batch;
        if (n_gpus > n_batches) n_gpus = n_batches;
        if (n_gpus > 64) n_gpus = 64;
    }
    feeder_t feeders[64];
    pthread_t th[64];
    for (int g = 0; g < n_gpus; g++) {
        feeder_t f = { st, &info, sel, n_sel, base, outdir, fmt, scale, device < 0 ? g : device,
                       threads / n_gpus > 0 ? threads / n_gpus : 1, batch, numbered, g, n_gpus, 0, 1 };
        feeders[g] = f;
    }
    int threaded[64] = {0};
    for (int g = 1; g < n_gpus; g++) threaded[g] = pthread_create(&th[g], NULL, feeder_main, &feeders[g]) == 0;
    feeder_main(&feeders[0]);
    int exported = 0, rc = 0;
    for (int g = 0; g < n_gpus; g++) {
        if (g > 0) {
            if (threaded[g]) pthread_join(th[g], NULL);
            else feeder_main(&feeders[g]);              /* no thread: take its share here */
        }
        exported += feeders[g].exported;
        if (feeders[g].rc) rc = 1;
    }
    if (n_exported) *n_exported = exported;
    mvf_close(st);
    free(sel);
    return rc ? MVG_FAILURE : MVG_SUCCESS;
}
