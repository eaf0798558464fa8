/*
 * mvgpu.h -- C ABI of the B200 intra-picture reconstruction path (libmvgpu.so).
 *
 * This header is the drop-in boundary for ONE path of MiniVideo: H.264 IDR
 * picture reconstruction + picture export.  File:line citations point into the
 * reference tree (minivideo/src/...), which is NOT part of this repository.
 *
 *   reference seam                                   replaced by
 *   ------------------------------------------------ -------------------------
 *   intra_prediction_process(dc, mb)                 mvg_upload()+mvg_run()  /
 *     decoder/h264/h264_intra_prediction.h:107,        mvg_decode_host()
 *     called once per MB at h264_macroblock.c:280     (once per BATCH of pictures)
 *   computeLevelScale4x4/8x8(dc, sps)                mvg_set_sps()
 *     decoder/h264/h264_transform.h:33-34,
 *     called at h264_parameterset.c:302-303
 *   export_idr(dc) -> export_idr_yuv420 / mb_to_rgb  mvg_download_yuv420(),
 *     export.c:618, export.c:65, export_utils.c:209    mvg_download_rgb()
 *
 * The reference calls its hot path per macroblock from inside the CAVLC parse
 * loop.  A GPU needs whole pictures, many at a time, so the seam moves to
 * "a batch of parsed pictures": the host front end (mvfront.h) fills the
 * structure-of-arrays below for every macroblock of every picture, and the
 * library reconstructs the batch.
 *
 * Conventions (same as the reference, typedef.h:40-42 / minivideo.h:89-149):
 *   every function returns MVG_SUCCESS (1), MVG_FAILURE (0) or
 *   MVG_UNSUPPORTED (-1); nothing ever calls exit()/abort(); there is NO CPU
 *   fallback -- without a CUDA device mvg_create() returns MVG_FAILURE.
 *   A context is not thread-safe; use one context per GPU per host thread.
 */
#ifndef MVGPU_H
#define MVGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVG_UNSUPPORTED (-1)
#define MVG_FAILURE       0
#define MVG_SUCCESS       1

/* mb_kind values == the reference's MbPartPredMode[0] for intra MBs
 * (h264_macroblock_struct.h: Intra_4x4, Intra_8x8, Intra_16x16). */
#define MVG_MB_I4x4   0
#define MVG_MB_I8x8   1
#define MVG_MB_I16x16 2

#define MVG_COEFF_PER_MB 384   /* 16x16 luma + 2 x 8x8 chroma levels */

/*
 * One batch of parsed pictures, structure-of-arrays.  All pictures of a batch
 * share the geometry/tables installed by the last mvg_set_sps().  With
 * N = width_mbs*height_mbs and P = n_pics every array is indexed
 * [pic*N + mbAddr] (mbAddr = raster macroblock address, h264_macroblock.c:375).
 *
 * Field <- reference Macroblock_t member (h264_macroblock_struct.h:209-319):
 *  mb_kind     <- MbPartPredMode[0]
 *  i16_mode    <- Intra16x16PredMode (0..3), 0 when mb_kind != I16x16
 *  chroma_mode <- IntraChromaPredMode (0 DC, 1 H, 2 V, 3 Plane)
 *  qp_y        <- QPY after the mb_qp_delta recurrence (h264_macroblock.c:263-269)
 *  cbp         <- CodedBlockPatternChroma<<4 | CodedBlockPatternLuma (a hint;
 *                 blocks whose bit is clear must still hold zero levels)
 *  luma_modes  <- FINAL Intra4x4PredMode[16] (mb_kind I4x4) or
 *                 Intra8x8PredMode[4] in entries 0..3 (I8x8); the host resolves
 *                 prev_intra*_pred_mode_flag / rem_intra*_pred_mode
 *                 (h264_intra_prediction.c:196-290, :977-1083).  Unused for I16x16.
 *  coeff       <- transform coefficient levels, int16, zig-zag order:
 *     [  0..255] luma:  I4x4  : LumaLevel4x4[blk][k]      at blk*16+k
 *                       I8x8  : LumaLevel8x8[blk8][k]     at blk8*64+k
 *                       I16x16: Intra16x16ACLevel[blk][k-1] at blk*16+k (k=1..15),
 *                               and at blk*16+0 the element c[i][j] of the
 *                               inverse-scanned Intra16x16DCLevel matrix
 *                               (h264_transform.c:180) with (i,j) the position of
 *                               blk in the MB (utils.h:54 raster_4x4_2d).
 *     [256..319] Cb:    ChromaDCLevel[0][blk] at 256+blk*16, ChromaACLevel[0][blk][k-1] at +k
 *     [320..383] Cr:    same with iCbCr = 1
 */
typedef struct mvg_batch {
    int32_t        n_pics;
    const uint8_t *mb_kind;      /* [P*N]      */
    const uint8_t *i16_mode;     /* [P*N]      */
    const uint8_t *chroma_mode;  /* [P*N]      */
    const int8_t  *qp_y;         /* [P*N]      */
    const uint8_t *cbp;          /* [P*N]      */
    const uint8_t *luma_modes;   /* [P*N*16]   */
    const int16_t *coeff;        /* [P*N*384]  */
} mvg_batch;

typedef struct mvg_ctx mvg_ctx;

/* Per-stage device timings of the last mvg_run(), CUDA events on the stream the
 * kernels were launched on (milliseconds). */
typedef struct mvg_timing {
    float k1_dequant_idct_ms;   /* kernel 1: dequant + inverse transforms        */
    float k2_wavefront_ms;      /* kernel 2: intra prediction + residual add     */
    float k3_rgb_ms;            /* kernel 3: 4:2:0 -> RGB24 (+ box downscale)    */
    float total_ms;             /* first launch -> last kernel end               */
    int32_t launches;           /* kernels launched by this run                  */
    float fused_ms;             /* fused pipeline: kf_recon (then k1 = k2 = 0)    */
} mvg_timing;

/* -- life cycle ----------------------------------------------------------- */

/* Create a context on CUDA device `device` able to hold `max_pics` pictures of
 * up to max_w_mbs x max_h_mbs macroblocks resident in HBM.
 * Replaces initDecodingContext()/decodeSPS() allocation (h264.c:208,
 * h264_parameterset.c:353).  FAILURE when no device / out of memory. */
int mvg_create(mvg_ctx **out, int device, int max_w_mbs, int max_h_mbs, int max_pics);
int mvg_destroy(mvg_ctx *ctx);

/* Last error message of this context (never NULL).  ctx may be NULL to read the
 * message of a failed mvg_create(). */
const char *mvg_last_error(const mvg_ctx *ctx);

/* Install picture geometry and dequantisation tables.
 *  level_scale4x4[c][q][i*4+j] = LevelScale4x4[c][q][i][j]  (c = Y,Cb,Cr)
 *  level_scale8x8[q][i*8+j]    = LevelScale8x8[0][q][i][j]
 * as built by computeLevelScale4x4/8x8 (h264_transform.c:645-741) from
 * normAdjust (h264.c:419-493) and the SPS scaling matrices.
 * cb/cr_qp_offset = chroma_qp_index_offset / second_chroma_qp_index_offset
 * (h264_transform.c:598-637). */
int mvg_set_sps(mvg_ctx *ctx, int width_mbs, int height_mbs,
                const int32_t level_scale4x4[3 * 6 * 16],
                const int32_t level_scale8x8[6 * 64],
                int cb_qp_offset, int cr_qp_offset);

/* Helper: build the two tables above from H.264 scaling lists in zig-zag order
 * (flat 16 when a pointer is NULL), i.e. what decodeSPS() +
 * computeLevelScale*() do (h264_parameterset.c:236-303).
 * lists4x4: [3][16] intra Y,Cb,Cr; list8x8: [64] intra Y. */
int mvg_build_level_scale(const uint8_t *lists4x4, const uint8_t *list8x8,
                          int32_t level_scale4x4[3 * 6 * 16],
                          int32_t level_scale8x8[6 * 64]);

/* -- resident path (inputs already in HBM when the timed region starts) ---- */

/* Copy a batch from HOST memory into the context's HBM input buffers
 * (synchronous).  Pictures land in slots first_slot .. first_slot+n_pics-1. */
int mvg_upload(mvg_ctx *ctx, const mvg_batch *host_batch, int first_slot);

/* Duplicate resident slot `src_slot` into `dst_slot` on the device (used by the
 * benchmark to fill HBM with more pictures than were generated on the host). */
int mvg_clone_slot(mvg_ctx *ctx, int src_slot, int dst_slot);

/* Reconstruct pictures [first_slot, first_slot+n_pics): kernel 1, kernel 2 and,
 * if rgb_scale >= 1, kernel 3 (RGB24 at 1/rgb_scale size; rgb_scale = 0 skips
 * it).  Asynchronous on the context stream; mvg_sync() waits. */
int mvg_run(mvg_ctx *ctx, int first_slot, int n_pics, int rgb_scale);
/* The same for callers that want the exported RGB picture and nothing else (what export_idr() does for the
 * bmp/tga/png formats, export.c:535-601): ONE kernel, levels in -> full-size RGB24 out, no intermediate picture
 * in HBM.  mvg_download_yuv420() is not available for these slots afterwards. */
int mvg_run_rgb(mvg_ctx *ctx, int first_slot, int n_pics);
/* RGB24 thumbnails and nothing else: the picture of mvg_run_rgb() averaged over rgb_scale x rgb_scale boxes (rounded;
 * no reference counterpart).  rgb_scale 2, 4, 8, 16: ONE kernel, levels in -> thumbnail
 * out (BASELINE.json configs[3]: "fused RGB thumbnail downscale"); rgb_scale 1 is mvg_run_rgb(); any other divisor of
 * the picture size runs the reconstruction to tiles and kernel 3.  mvg_download_rgb() fetches the result. */
int mvg_run_thumbs(mvg_ctx *ctx, int first_slot, int n_pics, int rgb_scale);
int mvg_sync(mvg_ctx *ctx);

/* Which kernels reconstruct: the fused kernel (default; dequantisation, transforms, prediction and -- for
 * mvg_run_rgb() and the RGB-only end-to-end calls -- the RGB conversion in one launch) or round 1's separate
 * kernels 1, 2, 3 (kept for comparison; same results).  The environment variable MVG_PIPELINE=split selects
 * the latter at mvg_create(). */
#define MVG_PIPELINE_FUSED 0
#define MVG_PIPELINE_SPLIT 1
int mvg_set_pipeline_mode(mvg_ctx *ctx, int mode);
int mvg_get_timing(mvg_ctx *ctx, mvg_timing *out);

/* Bracket a region of several mvg_run() calls with two CUDA events on the context
 * stream: mvg_mark(ctx, 0) before, mvg_mark(ctx, 1) after; mvg_mark_elapsed() waits
 * for the second event and returns the device time between them (milliseconds). */
int mvg_mark(mvg_ctx *ctx, int which);
int mvg_mark_elapsed(mvg_ctx *ctx, float *ms);

/* Copy results of picture slot `slot` to HOST memory.
 * yuv420: planar I420, (16*width_mbs) x (16*height_mbs), uncropped -- exactly the
 * bytes export_idr_yuv420() writes (export.c:100-151).
 * rgb: RGB24 interleaved, top-down, ((16*width_mbs)/scale) x ((16*height_mbs)/scale)
 * of the last mvg_run(); scale 1 is byte-identical to mb_to_rgb()
 * (export_utils.c:209-324). */
int mvg_download_yuv420(mvg_ctx *ctx, int slot, uint8_t *y, uint8_t *cb, uint8_t *cr);
int mvg_download_rgb(mvg_ctx *ctx, int slot, uint8_t *rgb);

/* Debug/verification tap: int16 residual of kernel 1 for one picture,
 * [N][384] = per MB 16x16 luma raster, 8x8 Cb raster, 8x8 Cr raster. */
int mvg_download_residual(mvg_ctx *ctx, int slot, int16_t *residual);

/* -- end-to-end path (host buffers in, host buffers out) ------------------- */

/* Reconstruct a batch given in HOST memory and return results to HOST memory:
 * H2D copies, the three kernels and D2H copies are pipelined over chunks of
 * pictures on separate streams.  Host buffers should be pinned
 * (mvg_host_alloc) for full PCIe speed.  yuv_out / rgb_out may be NULL.
 *   yuv_out: [P][1.5 * W * H]   rgb_out: [P][3 * (W/s) * (H/s)] */
int mvg_decode_host(mvg_ctx *ctx, const mvg_batch *host_batch,
                    uint8_t *yuv_out, uint8_t *rgb_out, int rgb_scale);

/* -- asynchronous form -----------------------------------------------------------
 * mvg_submit() enqueues everything mvg_decode_host() does -- H2D copies, kernels, D2H copies, over the same three
 * streams -- and returns without waiting for the device; mvg_wait() blocks until the outputs of that submission are
 * complete in host memory.  A caller overlaps its own work (parsing the next batch, encoding the previous one) with
 * the GPU without threads of its own, and may have several submissions in flight (at most 8): their chunks follow
 * each other through the context's slot regions in submission order.  The batch ARRAYS and the output buffers must
 * stay valid and untouched until mvg_wait() returns; the batch structure itself may go away after the call.
 * mvg_decode_host*() are mvg_submit*() + mvg_wait().  mvg_set_sps() drains the context before it changes tables. */
typedef int32_t mvg_ticket;
int mvg_submit(mvg_ctx *ctx, const mvg_batch *host_batch, uint8_t *yuv_out, uint8_t *rgb_out, int rgb_scale,
               mvg_ticket *ticket);
int mvg_wait(mvg_ctx *ctx, mvg_ticket ticket);
int mvg_poll(mvg_ctx *ctx, mvg_ticket ticket, int *done);     /* *done = 1 when mvg_wait() would not block */

/* -- packed transfer format ---------------------------------------------------
 * A parsed intra picture is mostly zero levels (a CAVLC block carries a handful of
 * them), and the end-to-end path is bound by PCIe, so the levels can cross the bus
 * packed.  The 384 levels of a macroblock are 24 chunks of 16 (chunk b =
 * coeff[16 b .. 16 b + 15] of mvg_batch, i.e. one residual_block_cavlc() call,
 * h264_cavlc.c:79, or a quarter of an 8x8 block):
 *   nz_blocks[mb]  bit b set <=> chunk b holds a non-zero level
 *   words          per macroblock: one uint16 mask per SET bit of nz_blocks in
 *                  ascending b (bit k set <=> level k of the chunk is non-zero),
 *                  followed by the non-zero levels (int16) of those chunks in the
 *                  same order, ascending k
 *   word_off[mb]   index (in uint16 units) of the macroblock's first word, counted
 *                  from the first word of its picture
 *   pic_off[i]     index of picture i's first word in `words`; pic_off[n_pics] =
 *                  total number of words
 * The other arrays are those of mvg_batch (cbp is not needed: nz_blocks says more).
 * A front end can emit this directly (mvf_parse_pictures_packed); mvg_pack_batch()
 * converts a dense batch.  On the device the levels are expanded to the dense
 * layout before kernel 1, so results are identical by construction. */
typedef struct mvg_packed_batch {
    int32_t         n_pics;
    const uint8_t  *mb_kind;      /* [P*N]    */
    const uint8_t  *i16_mode;     /* [P*N]    */
    const uint8_t  *chroma_mode;  /* [P*N]    */
    const int8_t   *qp_y;         /* [P*N]    */
    const uint8_t  *luma_modes;   /* [P*N*16] */
    const uint32_t *nz_blocks;    /* [P*N]    */
    const uint32_t *word_off;     /* [P*N]    */
    const uint64_t *pic_off;      /* [P+1]    */
    const uint16_t *words;        /* [pic_off[P]] */
} mvg_packed_batch;

/* Upper bound of the words one macroblock can need (24 masks + 384 levels). */
#define MVG_PACKED_WORDS_PER_MB 408

/* Pack `n_pics` pictures of `n_mbs` macroblocks each from dense levels (host code,
 * `n_threads` host threads; <= 0: one per core).  The caller provides nz_blocks[P*N],
 * word_off[P*N], pic_off[P+1] and words[words_capacity]; FAILURE when the capacity
 * is too small (P*N*MVG_PACKED_WORDS_PER_MB always suffices). */
int mvg_pack_batch(const int16_t *coeff, int n_pics, int n_mbs,
                   uint32_t *nz_blocks, uint32_t *word_off, uint64_t *pic_off,
                   uint16_t *words, size_t words_capacity, int n_threads);

/* mvg_decode_host() / mvg_submit() for a packed batch: same pipeline, same outputs. */
int mvg_decode_host_packed(mvg_ctx *ctx, const mvg_packed_batch *host_batch,
                           uint8_t *yuv_out, uint8_t *rgb_out, int rgb_scale);
int mvg_submit_packed(mvg_ctx *ctx, const mvg_packed_batch *host_batch,
                      uint8_t *yuv_out, uint8_t *rgb_out, int rgb_scale, mvg_ticket *ticket);

/* Pictures per pipeline chunk of mvg_decode_host() (0 = automatic: an eighth of the batch, at
 * most a third of the context).  Larger chunks fill the GPU better, smaller ones overlap the
 * PCIe copies of neighbouring chunks better. */
int mvg_set_pipeline(mvg_ctx *ctx, int chunk_pics);

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost). */
void *mvg_host_alloc(size_t bytes);
void  mvg_host_free(void *p);

/* Number of CUDA devices visible to the process (0 when there is none or the driver is missing). */
int mvg_device_count(void);

/* Geometry helpers. */
int mvg_width(const mvg_ctx *ctx);    /* luma width in samples  = 16*width_mbs  */
int mvg_height(const mvg_ctx *ctx);   /* luma height in samples = 16*height_mbs */
int mvg_max_pics(const mvg_ctx *ctx);
int mvg_sm_count(const mvg_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* MVGPU_H */
