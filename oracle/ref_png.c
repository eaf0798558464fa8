/*
 * ref_png.c -- TEST INFRASTRUCTURE: writes a raw RGB24 file as PNG through the reference's own vendored
 * stb_image_write (minivideo/src/stb_image_write.h, compiled into oracle/_ref/libminivideo_ref.a by export.c),
 * i.e. the exact call export_idr_png() makes (export.c:539).  Used by tests/test_png.py to pin mv_png.c.
 *   ref_png <w> <h> <in.rgb> <out> [png|bmp|tga]     (the same for stbi_write_bmp / stbi_write_tga, export.c:570,:601)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int stbi_write_png(char const *filename, int w, int h, int comp, const void *data, int stride_in_bytes);
int stbi_write_bmp(char const *filename, int w, int h, int comp, const void *data);
int stbi_write_tga(char const *filename, int w, int h, int comp, const void *data);

int main(int argc, char **argv)
{
    if (argc != 5 && argc != 6) { fprintf(stderr, "usage: ref_png <w> <h> <in.rgb> <out> [png|bmp|tga]\n"); return 2; }
    int w = atoi(argv[1]), h = atoi(argv[2]);
    size_t n = (size_t)w * h * 3;
    unsigned char *px = malloc(n ? n : 1);
    FILE *f = fopen(argv[3], "rb");
    if (!px || !f || fread(px, 1, n, f) != n) { fprintf(stderr, "ref_png: cannot read %s\n", argv[3]); return 1; }
    fclose(f);
    if (argc == 6 && !strcmp(argv[5], "bmp")) return stbi_write_bmp(argv[4], w, h, 3, px) ? 0 : 1;
    if (argc == 6 && !strcmp(argv[5], "tga")) return stbi_write_tga(argv[4], w, h, 3, px) ? 0 : 1;
    return stbi_write_png(argv[4], w, h, 3, px, w * 3) ? 0 : 1;
}
