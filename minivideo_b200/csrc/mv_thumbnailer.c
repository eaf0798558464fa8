/*
 * mv_thumbnailer.c -- command-line front of the GPU path, argument-compatible with the reference's
 * mini_thumbnailer (mini_thumbnailer/src/main.cpp:77-240):
 *
 *   mv_thumbnailer -i <file.264> [-o <dir>] [-f yuv420|bmp|tga] [-n N] [-e unfiltered|ordered|distributed]
 *                  [-s scale] [-d device] [-t threads] [-b batch]
 *
 * Annex-B file -> mvf_open_annexb() -> mvf_select_idr() (demuxer/filter.c semantics) ->
 * mvf_parse_pictures_packed() (threaded CAVLC, packed levels) -> mvg_decode_host_packed() (pinned copies +
 * kernels 0-4) -> picture
 * files named like export_idr() names them (export.c:627-642,:704-705): <input base name>[_<k>].<ext>
 * with k counting exported pictures when more than one was requested.  File contents are byte-identical
 * to the reference's: planar I420 (export.c:100-151), 24-bit bottom-up BMP and run-length TGA as
 * stb_image_write lays them out (export.c:535-539,:566-570).
 * Extensions over the reference: -o is honoured (the reference ignores it, h264.c:65), -s writes a
 * box-downscaled RGB thumbnail, -d/-t/-b pick the GPU, parser threads and batch size.
 * No CPU fallback: without a CUDA device the program fails.
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "mvfront.h"
#include "mvgpu.h"

enum { FMT_YUV420, FMT_BMP, FMT_TGA };

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

static void usage(void)
{
    fprintf(stderr, "usage: mv_thumbnailer -i <file.264> [-o <dir>] [-f yuv420|bmp|tga] [-n picture_number]\n"
                    "                      [-e unfiltered|ordered|distributed] [-s scale] [-d device] [-t threads] [-b batch]\n");
}

int main(int argc, char **argv)
{
    const char *in = NULL, *outdir = ".";
    int fmt = FMT_YUV420, n_want = 1, mode = 0, scale = 1, device = 0, threads = 0, batch = 64;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i], *v = i + 1 < argc ? argv[i + 1] : NULL;
        if (!strcmp(a, "-h") || !strcmp(a, "--help")) { usage(); return 0; }
        if (!v) { usage(); return 2; }
        if (!strcmp(a, "-i")) in = v;
        else if (!strcmp(a, "-o")) outdir = v;
        else if (!strcmp(a, "-f")) {
            if (!strcmp(v, "yuv420")) fmt = FMT_YUV420;
            else if (!strcmp(v, "bmp")) fmt = FMT_BMP;
            else if (!strcmp(v, "tga")) fmt = FMT_TGA;
            else { fprintf(stderr, "mv_thumbnailer: picture format '%s' is not supported (yuv420, bmp, tga)\n", v); return 2; }
        }
        else if (!strcmp(a, "-n")) n_want = atoi(v);
        else if (!strcmp(a, "-e")) {
            if (!strcmp(v, "unfiltered")) mode = 0;
            else if (!strcmp(v, "ordered")) mode = 1;
            else if (!strcmp(v, "distributed")) mode = 2;
            else { fprintf(stderr, "mv_thumbnailer: unknown extraction mode '%s'\n", v); return 2; }
        }
        else if (!strcmp(a, "-s")) scale = atoi(v);
        else if (!strcmp(a, "-d")) device = atoi(v);
        else if (!strcmp(a, "-t")) threads = atoi(v);
        else if (!strcmp(a, "-b")) batch = atoi(v);
        else if (!strcmp(a, "-q")) { /* JPEG quality: accepted, unused */ }
        else { fprintf(stderr, "mv_thumbnailer: unknown argument '%s'\n", a); usage(); return 2; }
        i++;
    }
    if (!in || n_want < 1 || scale < 1 || batch < 1) { usage(); return 2; }
    if (threads < 1) { long c = sysconf(_SC_NPROCESSORS_ONLN); threads = c > 0 ? (int)c : 1; }

    FILE *f = fopen(in, "rb");
    if (!f) { fprintf(stderr, "mv_thumbnailer: cannot open '%s'\n", in); return 1; }
    fseek(f, 0, SEEK_END);
    long flen = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *data = malloc((size_t)flen + 8);
    if (!data || fread(data, 1, (size_t)flen, f) != (size_t)flen) { fprintf(stderr, "mv_thumbnailer: read error\n"); return 1; }
    fclose(f);

    mvf_stream *st = NULL;
    if (mvf_open_annexb(data, (size_t)flen, &st) != MVG_SUCCESS) {
        fprintf(stderr, "mv_thumbnailer: %s\n", mvf_last_error(NULL));
        return 1;
    }
    mvf_info info;
    mvf_get_info(st, &info);
    int32_t *sel = malloc(sizeof(int32_t) * (size_t)(n_want > info.n_idr ? n_want : info.n_idr + 1));
    int n_sel = mvf_select_idr(st, n_want, mode, sel);
    if (n_sel < 1) { fprintf(stderr, "mv_thumbnailer: no picture to decode after filtering\n"); return 1; }
    /* the reference appends _<k> when more than one picture was requested after filtering (export.c:630) */
    int numbered;
    if (mode == 0) numbered = (n_want < info.n_idr ? n_want : info.n_idr) > 1;
    else {      /* picture_number after filtering = min(requested, candidates) = what 'ordered' would return */
        int32_t *tmp = malloc(sizeof(int32_t) * (size_t)(n_want > info.n_idr ? n_want : info.n_idr + 1));
        numbered = mvf_select_idr(st, n_want, 1, tmp) > 1;
        free(tmp);
    }
    const int W = 16 * info.width_mbs, H = 16 * info.height_mbs;
    if (fmt != FMT_YUV420 && (W % scale || H % scale)) { fprintf(stderr, "mv_thumbnailer: scale %d does not divide %dx%d\n", scale, W, H); return 2; }
    if (batch > n_sel) batch = n_sel;

    mvg_ctx *ctx = NULL;
    if (mvg_create(&ctx, device, info.width_mbs, info.height_mbs, batch) != MVG_SUCCESS) {
        fprintf(stderr, "mv_thumbnailer: %s\n", mvg_last_error(NULL));
        return 1;
    }
    if (mvg_set_sps(ctx, info.width_mbs, info.height_mbs, info.level_scale4x4, info.level_scale8x8,
                    info.cb_qp_offset, info.cr_qp_offset) != MVG_SUCCESS) {
        fprintf(stderr, "mv_thumbnailer: %s\n", mvg_last_error(ctx));
        return 1;
    }
    const size_t N = (size_t)info.width_mbs * info.height_mbs, nb = N * (size_t)batch;
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
    uint8_t *out = mvg_host_alloc((fmt == FMT_YUV420 ? yuv_sz : rgb_sz) * (size_t)batch);
    if (!pb.mb_kind || !pb.i16_mode || !pb.chroma_mode || !pb.qp_y || !pb.luma_modes || !pb.nz_blocks || !pb.word_off ||
        !pb.pic_off || !pb.words || !out) {
        fprintf(stderr, "mv_thumbnailer: pinned host allocation failed\n");
        return 1;
    }

    /* input base name without directory and extension (import.c: file_name) */
    char base[256];
    const char *slash = strrchr(in, '/');
    snprintf(base, sizeof base, "%s", slash ? slash + 1 : in);
    char *dot = strrchr(base, '.');
    if (dot) *dot = 0;
    static const char *ext[] = {"yuv", "bmp", "tga"};

    int exported = 0, rc = 0;
    for (int done = 0; done < n_sel && !rc; done += batch) {
        int cnt = n_sel - done < batch ? n_sel - done : batch;
        if (mvf_parse_pictures_packed(st, sel + done, 0, cnt, &pb, threads) != MVG_SUCCESS) {
            fprintf(stderr, "mv_thumbnailer: %s\n", mvf_last_error(st));
            rc = 1; break;
        }
        mvg_packed_batch gb = { cnt, pb.mb_kind, pb.i16_mode, pb.chroma_mode, pb.qp_y, pb.luma_modes,
                                pb.nz_blocks, pb.word_off, pb.pic_off, pb.words };
        int ok = fmt == FMT_YUV420 ? mvg_decode_host_packed(ctx, &gb, out, NULL, 0) : mvg_decode_host_packed(ctx, &gb, NULL, out, scale);
        if (ok != MVG_SUCCESS) { fprintf(stderr, "mv_thumbnailer: %s\n", mvg_last_error(ctx)); rc = 1; break; }
        for (int k = 0; k < cnt && !rc; k++) {
            char path[PATH_MAX];
            if (numbered) snprintf(path, sizeof path, "%s/%s_%d.%s", outdir, base, exported, ext[fmt]);
            else snprintf(path, sizeof path, "%s/%s.%s", outdir, base, ext[fmt]);
            int w = fmt == FMT_YUV420 ? write_raw(path, out + (size_t)k * yuv_sz, yuv_sz)
                  : fmt == FMT_BMP    ? write_bmp(path, out + (size_t)k * rgb_sz, ow, oh)
                                      : write_tga(path, out + (size_t)k * rgb_sz, ow, oh);
            if (!w) { fprintf(stderr, "mv_thumbnailer: cannot write '%s'\n", path); rc = 1; }
            else exported++;
        }
    }
    printf("mv_thumbnailer: %d picture(s) exported\n", exported);
    mvg_host_free(pb.mb_kind); mvg_host_free(pb.i16_mode); mvg_host_free(pb.chroma_mode); mvg_host_free(pb.qp_y);
    mvg_host_free(pb.luma_modes); mvg_host_free(pb.nz_blocks); mvg_host_free(pb.word_off); mvg_host_free(pb.pic_off);
    mvg_host_free(pb.words); mvg_host_free(out);
    mvg_destroy(ctx);
    mvf_close(st);
    free(sel); free(data);
    return rc;
}
