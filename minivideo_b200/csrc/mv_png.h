/*
 * mv_png.h -- RGB24 -> PNG, byte-identical to stb_image_write v1.01's stbi_write_png(w, h, 3, rgb, 3*w),
 * which is what the reference's export_idr_png() calls (minivideo/src/export.c:532-539).
 */
#ifndef MV_PNG_H
#define MV_PNG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Returns a malloc()ed PNG file image of the w x h RGB24 picture `rgb` (rows top-down, stride 3*w) and its
 * length in *out_len, or NULL (bad arguments / out of memory).  Thread-safe; free() the result. */
uint8_t *mvt_png_encode(const uint8_t *rgb, int w, int h, size_t *out_len);

#ifdef __cplusplus
}
#endif
#endif
