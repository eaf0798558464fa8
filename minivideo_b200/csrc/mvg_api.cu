/*
 * mvg_api.cu -- host side of libmvgpu.so: the C ABI of include/mvgpu.h.
 * Device memory, streams, table upload, kernel launches, pinned-copy pipeline.
 * There is no CPU fallback anywhere in this file: without a CUDA device every
 * entry point fails with MVG_FAILURE.
 */
#include "mvgpu.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>      /* header-only: the ranges cost two empty calls unless a profiler has injected itself */

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "mvg_internal.h"
#include "mvg_kernels.cuh"
#include "mvg_fused.cuh"

/* ------------------------------------------------------------------------- */

static char g_create_error[512] = "";

/* NVTX range around an entry point of the C ABI (tracing).  A timeline (Nsight Systems) then shows
 * the host calls above the copies and launches they enqueue. */
struct MvgRange {
    explicit MvgRange(const char *name) { nvtxRangePushA(name); }
    ~MvgRange() { nvtxRangePop(); }
};

#define MVG_PIPE_DEPTH 3       /* slot regions used by the end-to-end calls */
#define MVG_MAX_TICKETS 8      /* submissions in flight (mvg_submit*) */

/* one submission of the end-to-end path: completion event + the host staging of its picture offsets */
struct MvgTicket {
    cudaEvent_t done = nullptr;
    bool busy = false;
    uint64_t *h_picbase = nullptr;      /* pinned; read by asynchronous copies until `done` */
    int cap = 0;
};
#define MVG_WORK_RING  64      /* work counters, one per kernel-2 launch in flight */
#define MVG_K2_GROUP   1024    /* pictures interleaved row by row in kernel 2's claim order */

struct mvg_ctx {
    int device = -1, sm_count = 0;
    int k1_ctas_per_sm = 1, k2_ctas_per_sm = 1, kf_ctas_per_sm = 1;
    int mode = MVG_PIPELINE_FUSED;   /* mvg_set_pipeline_mode() */
    int stagger_override = -1;       /* DEV: row distance of the fused kernel from the environment */
    int max_w = 0, max_h = 0, max_pics = 0;
    int w_mbs = 0, h_mbs = 0;
    bool have_sps = false;
    char err[512] = "";

    cudaStream_t stream = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_mark[2] = {nullptr, nullptr};
    cudaEvent_t ev_h2d[MVG_PIPE_DEPTH] = {}, ev_comp[MVG_PIPE_DEPTH] = {}, ev_d2h[MVG_PIPE_DEPTH] = {};
    bool ran_k3 = false, ran_fused = false;
    bool tiles_valid = false;        /* the last run wrote macroblock tiles (planar YUV can be gathered) */
    int launches = 0, last_scale = 0;

    /* inputs (SoA), slot-major */
    uint8_t *d_kind = nullptr, *d_i16 = nullptr, *d_cm = nullptr, *d_cbp = nullptr, *d_modes = nullptr;
    int8_t *d_qp = nullptr;
    int16_t *d_coeff = nullptr;
    /* packed transfer format (allocated by the first mvg_decode_host_packed) */
    uint32_t *d_nzb = nullptr, *d_woff = nullptr;
    uint16_t *d_words = nullptr;
    uint64_t *d_picbase = nullptr;
    MvgTicket tickets[MVG_MAX_TICKETS];
    long long pipe_idx = 0;          /* chunks enqueued so far: region = pipe_idx % depth, across submissions */
    bool region_used[MVG_PIPE_DEPTH] = {};
    /* intermediates / outputs */
    int16_t *d_resid = nullptr;
    MvgMbCtl *d_ctl = nullptr;
    uint8_t *d_tiles = nullptr;      /* [slot][n_mb][384] reconstructed macroblocks (kernel 2 -> kernel 3 / 4) */
    uint8_t *d_yuv = nullptr, *d_rgb = nullptr;
    uint2 *d_halo = nullptr;         /* [slot][n_mb][8] flag-in-data bottom lines (kernel 2) */
    int *d_work = nullptr;
    unsigned long long *d_stats = nullptr;   /* DEV: kernel-2 cycle accounting (builds with -DMVG_K2_PROFILE) */
    int work_next = 0;
    unsigned epoch = 0;              /* bumped per kernel-2 launch; halo words carry it       */
    int pipe_chunk = 0;              /* pictures per mvg_decode_host() chunk, 0 = automatic    */
    MvgTables *d_tab = nullptr;
    MvgLuts *d_luts = nullptr;
    /* every device allocation (base pointer as returned by cudaMalloc, payload size, name); with MVG_DEBUG_GUARD=1 in the
     * environment each one sits between two guard bands filled with a pattern that mvg_debug_check() verifies, and its
     * payload starts out poisoned instead of zero -- the stand-in for compute-sanitizer's memcheck / initcheck */
    struct Alloc { void *base; size_t bytes; const char *name; };
    std::vector<Alloc> allocs;
    bool guard = false;

    size_t n_mb_max() const { return (size_t)max_w * max_h; }
    size_t n_mb() const { return (size_t)w_mbs * h_mbs; }
};

static int fail(mvg_ctx *ctx, const char *fmt, ...)
{
    char *dst = ctx ? ctx->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return MVG_FAILURE;
}

#define CK(ctx, call)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return fail(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

/* ------------------------------------------------------------------------- */
/* tables                                                                      */

static void zigzag(int n, uint8_t *zz)      /* zz[k] = row*n + col, spec 8.5.6 / 8.5.7 */
{
    int r = 0, c = 0;
    bool up = true;
    for (int k = 0; k < n * n; k++) {
        zz[k] = (uint8_t)(r * n + c);
        if (up) {
            if (c == n - 1) { r++; up = false; }
            else if (r == 0) { c++; up = false; }
            else { r--; c++; }
        } else {
            if (r == n - 1) { c++; up = true; }
            else if (c == 0) { r++; up = true; }
            else { r++; c--; }
        }
    }
}

extern "C" int mvg_build_level_scale(const uint8_t *lists4x4, const uint8_t *list8x8,
                                     int32_t ls4[3 * 6 * 16], int32_t ls8[6 * 64])
{
    /* normAdjust: h264.c:428-446, spec 8.5.9 */
    static const int v4[6][3] = {{10,16,13},{11,18,14},{13,20,16},{14,23,18},{16,25,20},{18,29,23}};
    static const int v8[6][6] = {{20,18,32,19,25,24},{22,19,35,21,28,26},{26,23,42,24,33,31},
                                 {28,25,45,26,35,33},{32,28,51,30,40,38},{36,32,58,34,46,43}};
    if (!ls4 || !ls8) return MVG_FAILURE;
    uint8_t zz4[16], zz8[64];
    zigzag(4, zz4); zigzag(8, zz8);
    for (int c = 0; c < 3; c++) {
        int m[16];
        for (int k = 0; k < 16; k++) m[zz4[k]] = lists4x4 ? lists4x4[c * 16 + k] : 16;
        for (int q = 0; q < 6; q++)
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 4; j++) {
                    const int cls = (i % 2 == 0 && j % 2 == 0) ? 0 : (i % 2 == 1 && j % 2 == 1) ? 1 : 2;
                    ls4[(c * 6 + q) * 16 + i * 4 + j] = m[i * 4 + j] * v4[q][cls];
                }
    }
    int m8[64];
    for (int k = 0; k < 64; k++) m8[zz8[k]] = list8x8 ? list8x8[k] : 16;
    for (int q = 0; q < 6; q++)
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++) {
                int cls;
                if (i % 4 == 0 && j % 4 == 0) cls = 0;
                else if (i % 2 == 1 && j % 2 == 1) cls = 1;
                else if (i % 4 == 2 && j % 4 == 2) cls = 2;
                else if ((i % 4 == 0 && j % 2 == 1) || (i % 2 == 1 && j % 4 == 0)) cls = 3;
                else if ((i % 4 == 0 && j % 4 == 2) || (i % 4 == 2 && j % 4 == 0)) cls = 4;
                else cls = 5;
                ls8[q * 64 + i * 8 + j] = m8[i * 8 + j] * v8[q][cls];
            }
    return MVG_SUCCESS;
}

/* A neighbour reference of a directional predictor: top row p[i,-1] (i = -1 is the
 * corner) or left column p[-1,i]. */
struct Ref { bool left; int i; };
static Ref T(int i) { return {false, i}; }
static Ref L(int i) { return i < 0 ? Ref{false, -1} : Ref{true, i}; }

struct Taps { Ref r[4]; int form; };                                    /* form: 1 copy, 2 two-tap, 3 three-tap, 4 end */
static Taps tap3(Ref p, Ref q, Ref r) { return {{p, q, q, r}, 3}; }   /* (p + 2q + r + 2) >> 2 */
static Taps tap2(Ref p, Ref q) { return {{p, p, q, q}, 2}; }          /* (p + q + 1) >> 1      */
static Taps tap1(Ref p) { return {{p, p, p, p}, 1}; }                 /* p                     */
static Taps tapend(Ref p, Ref q) { return {{p, q, q, q}, 4}; }        /* (p + 3q + 2) >> 2     */

/* spec 8.3.1.2.x (n = 4) and 8.3.2.2.x (n = 8); same formulas the reference codes at
 * h264_intra_prediction.c:496-926 and :1366-1793.  Mode 2 (DC) has no taps. */
static Taps nxn_taps(int n, int mode, int x, int y)
{
    switch (mode) {
    case 0: return tap1(T(x));
    case 1: return tap1(L(y));
    case 3: return (x == n - 1 && y == n - 1) ? tapend(T(2 * n - 2), T(2 * n - 1))
                                              : tap3(T(x + y), T(x + y + 1), T(x + y + 2));
    case 4:
        if (x > y) return tap3(T(x - y - 2), T(x - y - 1), T(x - y));
        if (x < y) return tap3(L(y - x - 2), L(y - x - 1), L(y - x));
        return tap3(T(0), T(-1), L(0));
    case 5: {
        const int z = 2 * x - y, i = x - (y >> 1);
        if (z >= 0 && !(z & 1)) return tap2(T(i - 1), T(i));
        if (z >= 0) return tap3(T(i - 2), T(i - 1), T(i));
        if (z == -1) return tap3(L(0), T(-1), T(0));
        return tap3(L(y - 2 * x - 1), L(y - 2 * x - 2), L(y - 2 * x - 3));
    }
    case 6: {
        const int z = 2 * y - x, i = y - (x >> 1);
        if (z >= 0 && !(z & 1)) return tap2(L(i - 1), L(i));
        if (z >= 0) return tap3(L(i - 2), L(i - 1), L(i));
        if (z == -1) return tap3(L(0), T(-1), T(0));
        return tap3(T(x - 2 * y - 1), T(x - 2 * y - 2), T(x - 2 * y - 3));
    }
    case 7: {
        const int i = x + (y >> 1);
        return (y & 1) ? tap3(T(i), T(i + 1), T(i + 2)) : tap2(T(i), T(i + 1));
    }
    case 8: {
        const int z = x + 2 * y, i = y + (x >> 1), zmax = 2 * n - 3;
        if (z > zmax) return tap1(L(n - 1));
        if (z == zmax) return tapend(L(n - 2), L(n - 1));
        return (z & 1) ? tap3(L(i), L(i + 1), L(i + 2)) : tap2(L(i), L(i + 1));
    }
    default: return tap1(T(0));
    }
}

extern "C" void mvg_build_luts(MvgLuts *out)
{
    memset(out, 0, sizeof *out);
    for (int row = 0; row < 16; row++) {
        const int mode = row == 11 ? 3 : row == 15 ? 7 : row;
        /* DC (h264_intra_prediction.c:554-600) as tap rows: row 9 = the four samples above, (sum + 2) >> 2;
         * row 10 = the four samples to the left; row 2 (both sides available) = even x the four above, odd x the
         * four to the left: kernel 2 adds the half-sum of the lane next door and shifts by 3 */
        const bool dc = row == 2 || row == 9 || row == 10;
        if ((mode > 8 && !dc)) continue;
        const bool tr = row < 9;
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++) {
                Taps t = nxn_taps(4, mode, x, y);
                if (dc) {
                    const bool above = row == 9 || (row == 2 && !(x & 1));
                    for (int k = 0; k < 4; k++) t.r[k] = above ? T(k) : L(k);
                }
                uint32_t word = 0;
                for (int k = 0; k < 4; k++) {
                    int off;
                    if (t.r[k].left) off = t.r[k].i * MVG_LT_STRIDE - 1;
                    else {
                        int i = t.r[k].i;
                        if (!tr && i > 3) i = 3;             /* h264_intra_prediction.c:431-439 */
                        off = -MVG_LT_STRIDE + i;
                    }
                    word |= (uint32_t)(off + MVG_LUT4_BIAS) << (8 * k);
                }
                out->lut4[row][y * 4 + x] = out->lut4[row][16 + y * 4 + x] = word;
            }
    }
    auto line8 = [](Ref r) { return r.left ? MVG_N8_LEFT(r.i) : MVG_N8_TOP(r.i); };
    for (int mode = 0; mode < 9; mode++)
        for (int lane = 0; lane < 32; lane++) {
            uint32_t word = 0;
            for (int s = 0; s < 2; s++) {
                const int x = ((lane >> 3) & 1) * 4 + (lane & 1) * 2 + s, y = (lane >> 4) * 4 + ((lane >> 1) & 3);
                int idx = MVG_N8_DC, variant = 0;
                if (mode != 2) {
                    const Taps t = nxn_taps(8, mode, x, y);
                    if (t.form == 1) { idx = line8(t.r[0]); variant = 0; }
                    else if (t.form == 2) {                       /* adjacent pair -> f2 of the lower index */
                        const int a = line8(t.r[0]), b = line8(t.r[3]);
                        idx = a < b ? a : b; variant = 1;
                    } else { idx = line8(t.r[1]); variant = 2; }  /* centre (or the line end of an end tap) */
                }
                word |= (uint32_t)(mode == 2 ? MVG_N8_DC : MVG_N8_IDX(idx, variant)) << (16 * s);
            }
            out->lut8[mode][lane] = word;
        }
}

/* ------------------------------------------------------------------------- */
/* life cycle                                                                  */

extern "C" const char *mvg_last_error(const mvg_ctx *ctx) { return ctx ? ctx->err : g_create_error; }

#define MVG_GUARD_BYTES ((size_t)64 << 10)
#define MVG_GUARD_BYTE  0xA5
#define MVG_POISON_BYTE 0x5A

static cudaError_t galloc(mvg_ctx *ctx, void **p, size_t bytes, const char *name)
{
    const size_t g = ctx->guard ? MVG_GUARD_BYTES : 0;
    void *base = nullptr;
    cudaError_t e = cudaMalloc(&base, bytes + 2 * g);
    if (e != cudaSuccess) return e;
    try { ctx->allocs.push_back({base, bytes, name}); } catch (...) { cudaFree(base); return cudaErrorMemoryAllocation; }
    if (ctx->guard) {
        e = cudaMemset(base, MVG_GUARD_BYTE, bytes + 2 * g);
        if (e == cudaSuccess) e = cudaMemset((uint8_t *)base + g, MVG_POISON_BYTE, bytes);
        if (e != cudaSuccess) return e;
    }
    *p = (uint8_t *)base + g;
    return cudaSuccess;
}
template <typename T>
static cudaError_t dalloc(mvg_ctx *ctx, T **p, size_t count, const char *name) { return galloc(ctx, (void **)p, count * sizeof(T), name); }

extern "C" int mvg_create(mvg_ctx **out, int device, int max_w_mbs, int max_h_mbs, int max_pics)
{
    if (!out) return fail(nullptr, "mvg_create: out is NULL");
    *out = nullptr;
    if (max_w_mbs < 1 || max_h_mbs < 1 || max_pics < 1) return fail(nullptr, "mvg_create: bad geometry");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, "mvg_create: no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) return fail(nullptr, "mvg_create: device %d out of range (0..%d)", device, n_dev - 1);

    mvg_ctx *ctx = new (std::nothrow) mvg_ctx;
    if (!ctx) return fail(nullptr, "mvg_create: out of host memory");
    ctx->device = device; ctx->max_w = max_w_mbs; ctx->max_h = max_h_mbs; ctx->max_pics = max_pics;
    if (const char *gd = getenv("MVG_DEBUG_GUARD")) ctx->guard = gd[0] == '1';

    auto bail = [&](const char *what, cudaError_t err) {
        fail(nullptr, "mvg_create: %s: %s", what, cudaGetErrorString(err));
        mvg_destroy(ctx);
        return MVG_FAILURE;
    };
#define TRY(what, call) do { cudaError_t e2 = (call); if (e2 != cudaSuccess) return bail(what, e2); } while (0)
    TRY("cudaSetDevice", cudaSetDevice(device));
    cudaDeviceProp prop;
    TRY("cudaGetDeviceProperties", cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    TRY("k1 shared memory", cudaFuncSetAttribute(k1_dequant_idct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(K1WarpSmem) * K1_WARPS)));
    TRY("occupancy k1", cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->k1_ctas_per_sm, k1_dequant_idct, K1_WARPS * 32, sizeof(K1WarpSmem) * K1_WARPS));
    TRY("k2 shared memory", cudaFuncSetAttribute(k2_wavefront, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K2_SMEM_BYTES));
    TRY("occupancy k2", cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->k2_ctas_per_sm, k2_wavefront, K2_WARPS * 32, K2_SMEM_BYTES));
    TRY("kf shared memory", cudaFuncSetAttribute(kf_recon<KF_OUT_TILES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KF_SMEM_BYTES(KF_OUT_TILES)));
    TRY("kf shared memory", cudaFuncSetAttribute(kf_recon<KF_OUT_RGB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KF_SMEM_BYTES(KF_OUT_RGB)));
    TRY("kf shared memory", cudaFuncSetAttribute(kf_recon<KF_OUT_RGBS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KF_SMEM_BYTES(KF_OUT_RGBS)));
    TRY("kf shared memory", cudaFuncSetAttribute(kf_recon<KF_OUT_RGBS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KF_SMEM_BYTES(KF_OUT_RGBS)));
    TRY("kf shared memory", cudaFuncSetAttribute(kf_recon<KF_OUT_RGBS, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KF_SMEM_BYTES(KF_OUT_RGBS)));
    TRY("kf shared memory", cudaFuncSetAttribute(kf_recon<KF_OUT_RGBS, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KF_SMEM_BYTES(KF_OUT_RGBS)));
    {
        int a = 0, b = 0;
        TRY("occupancy kf", cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, kf_recon<KF_OUT_RGB>, KF_WARPS_OF(KF_OUT_RGB) * 32, KF_SMEM_BYTES(KF_OUT_RGB)));
        TRY("occupancy kf", cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kf_recon<KF_OUT_TILES>, KF_WARPS_OF(KF_OUT_TILES) * 32, KF_SMEM_BYTES(KF_OUT_TILES)));
        ctx->kf_ctas_per_sm = std::min(a, b);
    }
    if (ctx->k1_ctas_per_sm < 1 || ctx->k2_ctas_per_sm < 1 || ctx->kf_ctas_per_sm < 1)
        return bail("kernel does not fit on an SM", cudaErrorLaunchOutOfResources);
    if (const char *m = getenv("MVG_PIPELINE")) ctx->mode = strcmp(m, "split") == 0 ? MVG_PIPELINE_SPLIT : MVG_PIPELINE_FUSED;
    if (const char *e = getenv("MVG_KF_STAGGER")) ctx->stagger_override = std::min(std::max(atoi(e), 0), 64);
    TRY("stream", cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    TRY("stream", cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    TRY("stream", cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    for (auto &ev : ctx->ev) TRY("event", cudaEventCreate(&ev));
    for (auto &ev : ctx->ev_mark) TRY("event", cudaEventCreate(&ev));
    for (int i = 0; i < MVG_PIPE_DEPTH; i++) {
        TRY("event", cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
        TRY("event", cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming));
        TRY("event", cudaEventCreateWithFlags(&ctx->ev_d2h[i], cudaEventDisableTiming));
    }
    const size_t n = ctx->n_mb_max() * (size_t)max_pics;
    TRY("alloc mb_kind", dalloc(ctx, &ctx->d_kind, n, "mb_kind"));
    TRY("alloc i16_mode", dalloc(ctx, &ctx->d_i16, n, "i16_mode"));
    TRY("alloc chroma_mode", dalloc(ctx, &ctx->d_cm, n, "chroma_mode"));
    TRY("alloc qp", dalloc(ctx, &ctx->d_qp, n, "qp"));
    TRY("alloc cbp", dalloc(ctx, &ctx->d_cbp, n, "cbp"));
    TRY("alloc luma_modes", dalloc(ctx, &ctx->d_modes, n * 16, "luma_modes"));
    TRY("alloc coeff", dalloc(ctx, &ctx->d_coeff, n * 384, "coeff"));
    /* d_resid, d_ctl (split pipeline / residual tap) and d_yuv (planar output) are allocated on first use */
    TRY("alloc tiles", dalloc(ctx, &ctx->d_tiles, n * 384, "tiles"));
    TRY("alloc rgb", dalloc(ctx, &ctx->d_rgb, n * 768, "rgb"));
    TRY("alloc halo", dalloc(ctx, &ctx->d_halo, n * 8, "halo"));
    if (!ctx->guard) TRY("clear halo", cudaMemset(ctx->d_halo, 0, n * 8 * sizeof(uint2)));      /* guard mode: poison is as good a "never written" value */
    TRY("alloc work", dalloc(ctx, &ctx->d_work, MVG_WORK_RING, "work"));
    TRY("alloc stats", dalloc(ctx, &ctx->d_stats, 16, "stats"));
    TRY("clear stats", cudaMemset(ctx->d_stats, 0, 16 * sizeof(unsigned long long)));
    TRY("alloc tables", dalloc(ctx, &ctx->d_tab, 1, "tables"));
    TRY("alloc luts", dalloc(ctx, &ctx->d_luts, 1, "luts"));
    MvgLuts luts;
    mvg_build_luts(&luts);
    TRY("upload luts", cudaMemcpy(ctx->d_luts, &luts, sizeof luts, cudaMemcpyHostToDevice));
#undef TRY
    *out = ctx;
    return MVG_SUCCESS;
}

extern "C" int mvg_destroy(mvg_ctx *ctx)
{
    if (!ctx) return MVG_FAILURE;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto &a : ctx->allocs) cudaFree(a.base);
    for (auto &t : ctx->tickets) { if (t.h_picbase) cudaFreeHost(t.h_picbase); if (t.done) cudaEventDestroy(t.done); }
    for (auto ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto ev : ctx->ev_mark) if (ev) cudaEventDestroy(ev);
    for (int i = 0; i < MVG_PIPE_DEPTH; i++) {
        if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
        if (ctx->ev_comp[i]) cudaEventDestroy(ctx->ev_comp[i]);
        if (ctx->ev_d2h[i]) cudaEventDestroy(ctx->ev_d2h[i]);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    delete ctx;
    return MVG_SUCCESS;
}

extern "C" int mvg_set_sps(mvg_ctx *ctx, int width_mbs, int height_mbs,
                           const int32_t level_scale4x4[3 * 6 * 16], const int32_t level_scale8x8[6 * 64],
                           int cb_qp_offset, int cr_qp_offset)
{
    MvgRange nvtx_range("mvg_set_sps");
    if (!ctx) return MVG_FAILURE;
    if (!level_scale4x4 || !level_scale8x8) return fail(ctx, "mvg_set_sps: NULL table");
    if (width_mbs < 1 || height_mbs < 1 || width_mbs > ctx->max_w || height_mbs > ctx->max_h)
        return fail(ctx, "mvg_set_sps: %dx%d MBs exceeds the context capacity %dx%d",
                    width_mbs, height_mbs, ctx->max_w, ctx->max_h);
    CK(ctx, cudaSetDevice(ctx->device));
    MvgTables t;
    memset(&t, 0, sizeof t);
    memcpy(t.ls4, level_scale4x4, sizeof t.ls4);
    memcpy(t.ls8, level_scale8x8, sizeof t.ls8);
    for (int c = 0; c < 3; c++)
        for (int qp = 0; qp < 52; qp++)
            for (int k = 0; k < 16; k++) {
                const int32_t v = t.ls4[c][qp % 6][k];
                t.ls4q[c][qp][k] = qp > 23 ? (int32_t)((uint32_t)v << (qp / 6 - 4)) : v;
            }
    uint8_t zz8[64];
    zigzag(8, zz8);
    for (int k = 0; k < 64; k++) t.zz8inv[zz8[k]] = (uint8_t)k;
    t.cb_qp_offset = cb_qp_offset; t.cr_qp_offset = cr_qp_offset;
    /* the tables are read by kernels that may still be queued (asynchronous submissions): drain first */
    CK(ctx, cudaStreamSynchronize(ctx->s_h2d));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->s_d2h));
    CK(ctx, cudaMemcpy(ctx->d_tab, &t, sizeof t, cudaMemcpyHostToDevice));
    ctx->w_mbs = width_mbs; ctx->h_mbs = height_mbs; ctx->have_sps = true;
    return MVG_SUCCESS;
}

extern "C" int mvg_set_pipeline(mvg_ctx *ctx, int chunk_pics)
{
    if (!ctx || chunk_pics < 0) return MVG_FAILURE;
    ctx->pipe_chunk = chunk_pics;
    return MVG_SUCCESS;
}

/* DEV: read and clear the kernel-2 cycle counters (all zero unless built with -DMVG_K2_PROFILE) */
extern "C" int mvg_dev_k2_stats(mvg_ctx *ctx, unsigned long long out[16])
{
    if (!ctx || !out) return MVG_FAILURE;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaMemcpy(out, ctx->d_stats, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CK(ctx, cudaMemset(ctx->d_stats, 0, 16 * sizeof(unsigned long long)));
    return MVG_SUCCESS;
}

/* DEV: with MVG_DEBUG_GUARD=1, count the guard-band bytes around every device allocation that no longer hold the
 * pattern (0 = nothing wrote outside its buffer); `report` receives the names of the damaged allocations */
extern "C" long long mvg_debug_check(mvg_ctx *ctx, char *report, size_t report_cap)
{
    if (!ctx) return -1;
    if (report && report_cap) report[0] = 0;
    if (!ctx->guard) return 0;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return -1;
    std::vector<uint8_t> h(MVG_GUARD_BYTES);
    long long bad = 0;
    for (auto &a : ctx->allocs)
        for (int side = 0; side < 2; side++) {
            const uint8_t *src = (const uint8_t *)a.base + (side ? MVG_GUARD_BYTES + a.bytes : 0);
            if (cudaMemcpy(h.data(), src, MVG_GUARD_BYTES, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
            long long here = 0;
            for (uint8_t b : h) here += b != MVG_GUARD_BYTE;
            if (here && report && report_cap) {
                const size_t o = strlen(report);
                snprintf(report + o, report_cap - o, "%s%s(%s): %lld bytes", o ? ", " : "", a.name, side ? "after" : "before", here);
            }
            bad += here;
        }
    return bad;
}

/* DEV: the violation counters of a -DMVG_CHECKED build ([0] shared-memory address outside the warp record, [1] list
 * index, [2] canary word overwritten, [3] table index); all zero -- and 0 returned -- in a product build */
extern "C" int mvg_debug_check_kernels(mvg_ctx *ctx, unsigned long long out[8])
{
    if (!ctx || !out) return MVG_FAILURE;
    memset(out, 0, 8 * sizeof(unsigned long long));
#ifdef MVG_CHECKED
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaDeviceSynchronize());
    CK(ctx, cudaMemcpyFromSymbol(out, mvg_check_fail, 8 * sizeof(unsigned long long)));
    return MVG_SUCCESS;
#else
    return MVG_UNSUPPORTED;
#endif
}

extern "C" int mvg_device_count(void)
{
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

extern "C" int mvg_width(const mvg_ctx *ctx) { return ctx ? 16 * ctx->w_mbs : 0; }
extern "C" int mvg_height(const mvg_ctx *ctx) { return ctx ? 16 * ctx->h_mbs : 0; }
extern "C" int mvg_max_pics(const mvg_ctx *ctx) { return ctx ? ctx->max_pics : 0; }
extern "C" int mvg_sm_count(const mvg_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" void *mvg_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void mvg_host_free(void *p) { if (p) cudaFreeHost(p); }

/* ------------------------------------------------------------------------- */
/* uploads                                                                     */

static int check_ready(mvg_ctx *ctx, int first_slot, int n_pics, const char *who)
{
    if (!ctx) return MVG_FAILURE;
    if (!ctx->have_sps) return fail(ctx, "%s: mvg_set_sps() has not been called", who);
    if (n_pics < 1 || first_slot < 0 || first_slot + n_pics > ctx->max_pics)
        return fail(ctx, "%s: slots [%d,%d) outside [0,%d)", who, first_slot, first_slot + n_pics, ctx->max_pics);
    return MVG_SUCCESS;
}

static int upload_async(mvg_ctx *ctx, const mvg_batch *b, int src_pic, int first_slot, int n_pics, cudaStream_t st)
{
    const size_t n = ctx->n_mb(), o = (size_t)first_slot * n, s = (size_t)src_pic * n, cnt = (size_t)n_pics * n;
    CK(ctx, cudaMemcpyAsync(ctx->d_kind + o, b->mb_kind + s, cnt, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(ctx->d_i16 + o, b->i16_mode + s, cnt, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(ctx->d_cm + o, b->chroma_mode + s, cnt, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(ctx->d_qp + o, b->qp_y + s, cnt, cudaMemcpyHostToDevice, st));
    if (b->cbp) CK(ctx, cudaMemcpyAsync(ctx->d_cbp + o, b->cbp + s, cnt, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(ctx->d_modes + o * 16, b->luma_modes + s * 16, cnt * 16, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(ctx->d_coeff + o * 384, b->coeff + s * 384, cnt * 768, cudaMemcpyHostToDevice, st));
    return MVG_SUCCESS;
}

static int check_batch(mvg_ctx *ctx, const mvg_batch *b, const char *who)
{
    if (!b) return fail(ctx, "%s: batch is NULL", who);
    if (!b->mb_kind || !b->i16_mode || !b->chroma_mode || !b->qp_y || !b->luma_modes || !b->coeff)
        return fail(ctx, "%s: a required SoA pointer is NULL", who);
    return MVG_SUCCESS;
}

extern "C" int mvg_upload(mvg_ctx *ctx, const mvg_batch *b, int first_slot)
{
    MvgRange nvtx_range("mvg_upload");
    if (!ctx) return MVG_FAILURE;
    if (check_batch(ctx, b, "mvg_upload") != MVG_SUCCESS) return MVG_FAILURE;
    if (check_ready(ctx, first_slot, b->n_pics, "mvg_upload") != MVG_SUCCESS) return MVG_FAILURE;
    CK(ctx, cudaSetDevice(ctx->device));
    if (upload_async(ctx, b, 0, first_slot, b->n_pics, ctx->stream) != MVG_SUCCESS) return MVG_FAILURE;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return MVG_SUCCESS;
}

extern "C" int mvg_clone_slot(mvg_ctx *ctx, int src_slot, int dst_slot)
{
    MvgRange nvtx_range("mvg_clone_slot");
    if (check_ready(ctx, src_slot, 1, "mvg_clone_slot") != MVG_SUCCESS) return MVG_FAILURE;
    if (check_ready(ctx, dst_slot, 1, "mvg_clone_slot") != MVG_SUCCESS) return MVG_FAILURE;
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t n = ctx->n_mb(), s = (size_t)src_slot * n, d = (size_t)dst_slot * n;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemcpyAsync(ctx->d_kind + d, ctx->d_kind + s, n, cudaMemcpyDeviceToDevice, st));
    CK(ctx, cudaMemcpyAsync(ctx->d_i16 + d, ctx->d_i16 + s, n, cudaMemcpyDeviceToDevice, st));
    CK(ctx, cudaMemcpyAsync(ctx->d_cm + d, ctx->d_cm + s, n, cudaMemcpyDeviceToDevice, st));
    CK(ctx, cudaMemcpyAsync(ctx->d_qp + d, ctx->d_qp + s, n, cudaMemcpyDeviceToDevice, st));
    CK(ctx, cudaMemcpyAsync(ctx->d_cbp + d, ctx->d_cbp + s, n, cudaMemcpyDeviceToDevice, st));
    CK(ctx, cudaMemcpyAsync(ctx->d_modes + d * 16, ctx->d_modes + s * 16, n * 16, cudaMemcpyDeviceToDevice, st));
    CK(ctx, cudaMemcpyAsync(ctx->d_coeff + d * 384, ctx->d_coeff + s * 384, n * 768, cudaMemcpyDeviceToDevice, st));
    return MVG_SUCCESS;
}

/* ------------------------------------------------------------------------- */
/* launches                                                                    */

/* buffers only some paths need */
static int ensure_yuv(mvg_ctx *ctx)
{
    if (!ctx->d_yuv) CK(ctx, dalloc(ctx, &ctx->d_yuv, ctx->n_mb_max() * (size_t)ctx->max_pics * 384, "yuv"));
    return MVG_SUCCESS;
}
static int ensure_split(mvg_ctx *ctx)
{
    const size_t n = ctx->n_mb_max() * (size_t)ctx->max_pics;
    if (!ctx->d_resid) CK(ctx, dalloc(ctx, &ctx->d_resid, n * 384, "residual"));
    if (!ctx->d_ctl) CK(ctx, dalloc(ctx, &ctx->d_ctl, n, "ctl"));
    return MVG_SUCCESS;
}

/* planar I420 of slots [first_slot, first_slot + n_pics) from the macroblock tiles (kernel 4) */
static int launch_planar(mvg_ctx *ctx, int first_slot, int n_pics, cudaStream_t st)
{
    if (ensure_yuv(ctx) != MVG_SUCCESS) return MVG_FAILURE;
    K3Params p;
    p.tiles = ctx->d_tiles; p.yuv = ctx->d_yuv; p.rgb = nullptr; p.width = 16 * ctx->w_mbs; p.height = 16 * ctx->h_mbs;
    p.scale = 1; p.first_slot = first_slot; p.n_pics = n_pics;
    const long long threads = (long long)ctx->n_mb() * 12 * n_pics;
    const int grid = (int)std::min<long long>((threads + 255) / 256, (long long)ctx->sm_count * 32);
    k4_yuv_planar<<<grid, 256, 0, st>>>(p);
    CK(ctx, cudaGetLastError());
    return MVG_SUCCESS;
}

/* kernel 1 alone over slots [first_slot, first_slot + n_pics): residual + control records (split pipeline, residual tap) */
static int launch_k1(mvg_ctx *ctx, int first_slot, int n_pics, cudaStream_t st)
{
    if (ensure_split(ctx) != MVG_SUCCESS) return MVG_FAILURE;
    const size_t n = ctx->n_mb();
    K1Params p;
    const size_t o = (size_t)first_slot * n;
    p.mb_kind = ctx->d_kind + o; p.i16_mode = ctx->d_i16 + o; p.chroma_mode = ctx->d_cm + o;
    p.luma_modes = ctx->d_modes + o * 16; p.qp_y = ctx->d_qp + o; p.coeff = ctx->d_coeff + o * 384;
    p.resid = ctx->d_resid + o * 384; p.ctl = ctx->d_ctl + o; p.tab = ctx->d_tab;
    p.n_mbs = (long long)n * n_pics;
    const long long groups = (p.n_mbs + K1_GROUP - 1) / K1_GROUP;
    const long long want = (groups + K1_WARPS - 1) / K1_WARPS;
    const int grid = (int)std::min<long long>(want, (long long)ctx->sm_count * ctx->k1_ctas_per_sm);
    k1_dequant_idct<<<grid, K1_WARPS * 32, sizeof(K1WarpSmem) * K1_WARPS, st>>>(p);
    CK(ctx, cudaGetLastError());
    return MVG_SUCCESS;
}

static int launch_k3(mvg_ctx *ctx, int first_slot, int n_pics, int rgb_scale, cudaStream_t st)
{
    const size_t n = ctx->n_mb();
    const int width = 16 * ctx->w_mbs, height = 16 * ctx->h_mbs;
    K3Params p;
    p.tiles = ctx->d_tiles; p.yuv = nullptr; p.rgb = ctx->d_rgb; p.width = width; p.height = height; p.scale = rgb_scale;
    p.first_slot = first_slot; p.n_pics = n_pics;
    const long long threads = rgb_scale == 1 ? (long long)n * 8 * n_pics       /* 8 lanes per macroblock */
                                             : (long long)(width / rgb_scale) * (height / rgb_scale) * n_pics;
    const int grid = (int)std::min<long long>((threads + 255) / 256, (long long)ctx->sm_count * 32);
    const bool pow2 = rgb_scale == 2 || rgb_scale == 4 || rgb_scale == 8 || rgb_scale == 16;
    if (rgb_scale == 1) k3_rgb_full<<<grid, 256, 0, st>>>(p);
    else if (pow2) k3_rgb_scaled<<<(int)std::min<long long>(((long long)n * n_pics + 31) / 32, (long long)ctx->sm_count * 16), 256, 0, st>>>(p);
    else k3_rgb_scaled_generic<<<grid, 256, 0, st>>>(p);
    CK(ctx, cudaGetLastError());
    return MVG_SUCCESS;
}

/* The reconstruction stages for slots [first_slot, first_slot + n_pics).
 *   fused pipeline (default): kf_recon<RGB> alone when only full-size RGB24 is wanted (`want_tiles` false and
 *     rgb_scale 1); otherwise kf_recon<TILES> and, for rgb_scale >= 1, kernel 3 on the tiles;
 *   split pipeline (mvg_set_pipeline_mode): kernel 1, kernel 2, kernel 3 as in round 1. */
static int launch_stages(mvg_ctx *ctx, int first_slot, int n_pics, int rgb_scale, bool want_tiles, cudaStream_t st, bool timed)
{
    const int W = ctx->w_mbs, H = ctx->h_mbs;
    const size_t n = ctx->n_mb();
    const int width = 16 * W, height = 16 * H;
    if (rgb_scale < 0 || (rgb_scale > 0 && (width % rgb_scale || height % rgb_scale)))
        return fail(ctx, "rgb_scale %d does not divide %dx%d", rgb_scale, width, height);

    int *work = ctx->d_work + ctx->work_next;
    ctx->work_next = (ctx->work_next + 1) % MVG_WORK_RING;
    CK(ctx, cudaMemsetAsync(work, 0, sizeof(int), st));
    if (++ctx->epoch == 0) ctx->epoch = 1;      /* 0 is the value of never-written words */

    const bool fused = ctx->mode == MVG_PIPELINE_FUSED;
    /* RGB-only requests go through one kernel: full size, or a box downscale by 2, 4, 8, 16 (other factors: tiles + kernel 3) */
    const bool thumbs = rgb_scale == 2 || rgb_scale == 4 || rgb_scale == 8 || rgb_scale == 16;
    const bool rgb_direct = fused && !want_tiles && (rgb_scale == 1 || thumbs);
    int launches = 0;
    if (timed) CK(ctx, cudaEventRecord(ctx->ev[0], st));
    if (!fused) {
        if (launch_k1(ctx, first_slot, n_pics, st) != MVG_SUCCESS) return MVG_FAILURE;
        launches++;
    }
    if (timed) CK(ctx, cudaEventRecord(ctx->ev[1], st));
    const long long items = (long long)n_pics * H;
    if (fused) {
        KFParams p;
        p.mb_kind = ctx->d_kind; p.i16_mode = ctx->d_i16; p.chroma_mode = ctx->d_cm; p.luma_modes = ctx->d_modes;
        p.qp_y = ctx->d_qp; p.coeff = ctx->d_coeff; p.tiles = ctx->d_tiles; p.rgb = ctx->d_rgb; p.halo = ctx->d_halo;
        p.work = work; p.tab = ctx->d_tab; p.luts = ctx->d_luts; p.epoch = ctx->epoch;
        p.w_mbs = W; p.h_mbs = H; p.first_slot = first_slot; p.n_pics = n_pics; p.group = MVG_K2_GROUP;
        p.sel[0] = 1u; p.sel[1] = 1u << 8; p.sel[2] = 1u << 16; p.sel[3] = 1u << 24;
        p.stats = ctx->d_stats;
        /* A launch with fewer rows than twice the GPU's warps is bound by the critical path through a picture (row r
         * starts `stagger` macroblocks behind row r - 1), not by throughput: let its rows follow each other as closely as the
         * group rule allows.  Large launches keep the distance that prevents rows from running in lock step. */
        p.stagger = items < 2LL * ctx->sm_count * KF_WARPS ? KF_GROUP + 1 : KF_STAGGER;
        if (ctx->stagger_override >= 0) p.stagger = ctx->stagger_override;      /* DEV: MVG_KF_STAGGER, read once at mvg_create() */
        /* one CTA per SM; a small batch is spread over as many SMs as it has rows (warps without a row exit at once):
         * a row's warp then has a scheduler to itself instead of sharing it with five others */
        const int grid = (int)std::min<long long>(items, (long long)ctx->sm_count * ctx->kf_ctas_per_sm);
        if (rgb_direct && thumbs) {
            const unsigned bt = KF_WARPS_OF(KF_OUT_RGBS) * 32;
            const size_t sm = KF_SMEM_BYTES(KF_OUT_RGBS);
            if (rgb_scale == 2)      kf_recon<KF_OUT_RGBS, 1><<<grid, bt, sm, st>>>(p);
            else if (rgb_scale == 4) kf_recon<KF_OUT_RGBS, 2><<<grid, bt, sm, st>>>(p);
            else if (rgb_scale == 8) kf_recon<KF_OUT_RGBS, 3><<<grid, bt, sm, st>>>(p);
            else                     kf_recon<KF_OUT_RGBS, 4><<<grid, bt, sm, st>>>(p);
        } else if (rgb_direct) kf_recon<KF_OUT_RGB><<<grid, KF_WARPS_OF(KF_OUT_RGB) * 32, KF_SMEM_BYTES(KF_OUT_RGB), st>>>(p);
        else kf_recon<KF_OUT_TILES><<<grid, KF_WARPS_OF(KF_OUT_TILES) * 32, KF_SMEM_BYTES(KF_OUT_TILES), st>>>(p);
        launches++;
    } else {
        K2Params p;
        p.resid = ctx->d_resid; p.ctl = ctx->d_ctl; p.tiles = ctx->d_tiles; p.halo = ctx->d_halo;
        p.epoch = ctx->epoch; p.group = MVG_K2_GROUP; p.stats = ctx->d_stats;
        p.sel[0] = 1u; p.sel[1] = 1u << 8; p.sel[2] = 1u << 16; p.sel[3] = 1u << 24;
        p.work = work; p.luts = ctx->d_luts; p.w_mbs = W; p.h_mbs = H; p.first_slot = first_slot; p.n_pics = n_pics;
        const int grid = (int)std::min<long long>((items + K2_WARPS - 1) / K2_WARPS, (long long)ctx->sm_count * ctx->k2_ctas_per_sm);
        k2_wavefront<<<grid, K2_WARPS * 32, K2_SMEM_BYTES, st>>>(p);
        launches++;
    }
    CK(ctx, cudaGetLastError());
    if (timed) CK(ctx, cudaEventRecord(ctx->ev[2], st));
    const bool run_k3 = rgb_scale >= 1 && !rgb_direct;
    if (run_k3) {
        if (launch_k3(ctx, first_slot, n_pics, rgb_scale, st) != MVG_SUCCESS) return MVG_FAILURE;
        launches++;
    }
    if (timed) {
        CK(ctx, cudaEventRecord(ctx->ev[3], st));
        ctx->ran_k3 = run_k3;
        ctx->ran_fused = fused;
        ctx->launches = launches;
    }
    ctx->last_scale = rgb_scale;
    ctx->tiles_valid = !rgb_direct;
    return MVG_SUCCESS;
}

extern "C" int mvg_run(mvg_ctx *ctx, int first_slot, int n_pics, int rgb_scale)
{
    MvgRange nvtx_range("mvg_run");
    if (check_ready(ctx, first_slot, n_pics, "mvg_run") != MVG_SUCCESS) return MVG_FAILURE;
    CK(ctx, cudaSetDevice(ctx->device));
    return launch_stages(ctx, first_slot, n_pics, rgb_scale, true, ctx->stream, true);
}

extern "C" int mvg_run_thumbs(mvg_ctx *ctx, int first_slot, int n_pics, int rgb_scale)
{
    MvgRange nvtx_range("mvg_run_thumbs");
    if (check_ready(ctx, first_slot, n_pics, "mvg_run_thumbs") != MVG_SUCCESS) return MVG_FAILURE;
    if (rgb_scale < 1) return fail(ctx, "mvg_run_thumbs: rgb_scale %d", rgb_scale);
    CK(ctx, cudaSetDevice(ctx->device));
    return launch_stages(ctx, first_slot, n_pics, rgb_scale, false, ctx->stream, true);
}

extern "C" int mvg_run_rgb(mvg_ctx *ctx, int first_slot, int n_pics)
{
    MvgRange nvtx_range("mvg_run_rgb");
    if (check_ready(ctx, first_slot, n_pics, "mvg_run_rgb") != MVG_SUCCESS) return MVG_FAILURE;
    CK(ctx, cudaSetDevice(ctx->device));
    return launch_stages(ctx, first_slot, n_pics, 1, false, ctx->stream, true);
}

extern "C" int mvg_set_pipeline_mode(mvg_ctx *ctx, int mode)
{
    if (!ctx || (mode != MVG_PIPELINE_FUSED && mode != MVG_PIPELINE_SPLIT)) return MVG_FAILURE;
    ctx->mode = mode;
    return MVG_SUCCESS;
}

extern "C" int mvg_sync(mvg_ctx *ctx)
{
    if (!ctx) return MVG_FAILURE;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->s_h2d));
    CK(ctx, cudaStreamSynchronize(ctx->s_d2h));
    return MVG_SUCCESS;
}

extern "C" int mvg_get_timing(mvg_ctx *ctx, mvg_timing *out)
{
    if (!ctx || !out) return MVG_FAILURE;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaEventSynchronize(ctx->ev[3]));
    out->k1_dequant_idct_ms = out->k2_wavefront_ms = out->fused_ms = 0.f;
    if (ctx->ran_fused) CK(ctx, cudaEventElapsedTime(&out->fused_ms, ctx->ev[1], ctx->ev[2]));
    else {
        CK(ctx, cudaEventElapsedTime(&out->k1_dequant_idct_ms, ctx->ev[0], ctx->ev[1]));
        CK(ctx, cudaEventElapsedTime(&out->k2_wavefront_ms, ctx->ev[1], ctx->ev[2]));
    }
    out->k3_rgb_ms = 0.f;
    if (ctx->ran_k3) CK(ctx, cudaEventElapsedTime(&out->k3_rgb_ms, ctx->ev[2], ctx->ev[3]));
    CK(ctx, cudaEventElapsedTime(&out->total_ms, ctx->ev[0], ctx->ev[3]));
    out->launches = ctx->launches;
    return MVG_SUCCESS;
}

extern "C" int mvg_mark(mvg_ctx *ctx, int which)
{
    if (!ctx || which < 0 || which > 1) return MVG_FAILURE;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaEventRecord(ctx->ev_mark[which], ctx->stream));
    return MVG_SUCCESS;
}

extern "C" int mvg_mark_elapsed(mvg_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return MVG_FAILURE;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaEventSynchronize(ctx->ev_mark[1]));
    CK(ctx, cudaEventElapsedTime(ms, ctx->ev_mark[0], ctx->ev_mark[1]));
    return MVG_SUCCESS;
}

/* ------------------------------------------------------------------------- */
/* downloads                                                                   */

static size_t rgb_bytes(const mvg_ctx *ctx, int scale)
{
    return (size_t)(16 * ctx->w_mbs / scale) * (16 * ctx->h_mbs / scale) * 3;
}

extern "C" int mvg_download_yuv420(mvg_ctx *ctx, int slot, uint8_t *y, uint8_t *cb, uint8_t *cr)
{
    MvgRange nvtx_range("mvg_download_yuv420");
    if (check_ready(ctx, slot, 1, "mvg_download_yuv420") != MVG_SUCCESS) return MVG_FAILURE;
    if (!ctx->tiles_valid) return fail(ctx, "mvg_download_yuv420: the last run produced RGB24 only (mvg_run_rgb, mvg_run_thumbs); use mvg_run()");
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t n = ctx->n_mb();
    if (launch_planar(ctx, slot, 1, ctx->stream) != MVG_SUCCESS) return MVG_FAILURE;
    const uint8_t *src = ctx->d_yuv + (size_t)slot * n * 384;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (y) CK(ctx, cudaMemcpy(y, src, n * 256, cudaMemcpyDeviceToHost));
    if (cb) CK(ctx, cudaMemcpy(cb, src + n * 256, n * 64, cudaMemcpyDeviceToHost));
    if (cr) CK(ctx, cudaMemcpy(cr, src + n * 320, n * 64, cudaMemcpyDeviceToHost));
    return MVG_SUCCESS;
}

extern "C" int mvg_download_rgb(mvg_ctx *ctx, int slot, uint8_t *rgb)
{
    MvgRange nvtx_range("mvg_download_rgb");
    if (check_ready(ctx, slot, 1, "mvg_download_rgb") != MVG_SUCCESS) return MVG_FAILURE;
    if (!rgb || ctx->last_scale < 1) return fail(ctx, "mvg_download_rgb: no RGB output (last run had rgb_scale 0)");
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t sz = rgb_bytes(ctx, ctx->last_scale);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaMemcpy(rgb, ctx->d_rgb + (size_t)slot * sz, sz, cudaMemcpyDeviceToHost));
    return MVG_SUCCESS;
}

extern "C" int mvg_download_residual(mvg_ctx *ctx, int slot, int16_t *residual)
{
    MvgRange nvtx_range("mvg_download_residual");
    if (check_ready(ctx, slot, 1, "mvg_download_residual") != MVG_SUCCESS) return MVG_FAILURE;
    if (!residual) return fail(ctx, "mvg_download_residual: NULL");
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t n = ctx->n_mb();
    /* the fused kernel never writes the residual to HBM: run kernel 1 (the same transform code) on this slot */
    if (launch_k1(ctx, slot, 1, ctx->stream) != MVG_SUCCESS) return MVG_FAILURE;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    int16_t *tmp = new (std::nothrow) int16_t[n * 384];
    if (!tmp) return fail(ctx, "mvg_download_residual: out of host memory");
    cudaError_t e = cudaMemcpy(tmp, ctx->d_resid + (size_t)slot * n * 384, n * 768, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { delete[] tmp; return fail(ctx, "cudaMemcpy failed: %s", cudaGetErrorString(e)); }
    /* device layout is block-major (4x4 block b at [b*16, b*16+16)); the ABI promises rasters */
    for (size_t mb = 0; mb < n; mb++) {
        const int16_t *src = tmp + mb * 384;
        int16_t *dst = residual + mb * 384;
        for (int b = 0; b < 16; b++) {
            const int bx = (b & 1) | (((b >> 2) & 1) << 1), by = ((b >> 1) & 1) | ((b >> 3) << 1);
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 4; j++) dst[(by * 4 + i) * 16 + bx * 4 + j] = src[b * 16 + i * 4 + j];
        }
        for (int p = 0; p < 2; p++)
            for (int b = 0; b < 4; b++)
                for (int i = 0; i < 4; i++)
                    for (int j = 0; j < 4; j++)
                        dst[256 + p * 64 + ((b >> 1) * 4 + i) * 8 + (b & 1) * 4 + j] = src[256 + p * 64 + b * 16 + i * 4 + j];
    }
    delete[] tmp;
    return MVG_SUCCESS;
}

/* ------------------------------------------------------------------------- */
/* end-to-end: host SoA in, host pictures out, pipelined over slot regions     */

static int take_ticket(mvg_ctx *ctx, mvg_ticket *out)
{
    for (int i = 0; i < MVG_MAX_TICKETS; i++)
        if (!ctx->tickets[i].busy) {
            /* blocking sync: a caller waiting in mvg_wait() sleeps instead of spinning on a host core the parser could use */
            if (!ctx->tickets[i].done) CK(ctx, cudaEventCreateWithFlags(&ctx->tickets[i].done, cudaEventDisableTiming | cudaEventBlockingSync));
            *out = i;
            return MVG_SUCCESS;
        }
    return fail(ctx, "too many submissions in flight (%d): mvg_wait() for one first", MVG_MAX_TICKETS);
}

/* chunk loop shared by the dense and the packed entry points: `upload(done, slot0, cnt)` enqueues the H2D
 * copies of pictures [done, done+cnt) into slots [slot0, ..) on ctx->s_h2d, `expand(slot0, cnt)` enqueues
 * whatever must run on ctx->stream before the reconstruction.  Nothing here waits for the device: the chunks of
 * one submission and of the next one follow each other through the same MVG_PIPE_DEPTH slot regions, ordered by
 * events (a region's inputs are overwritten after the compute that read them, its outputs after the D2H copy
 * that read them).  Completion = the event recorded behind the last D2H copy. */
template <typename Upload, typename Expand>
static int decode_pipeline(mvg_ctx *ctx, int n_pics, uint8_t *yuv_out, uint8_t *rgb_out, int rgb_scale, cudaEvent_t done_ev,
                           Upload upload, Expand expand)
{
    const int scale = rgb_out ? rgb_scale : 0;
    const size_t n = ctx->n_mb();
    const size_t yuv_sz = n * 384, rgb_sz = scale ? rgb_bytes(ctx, scale) : 0;
    if (yuv_out && ensure_yuv(ctx) != MVG_SUCCESS) return MVG_FAILURE;
    /* slot regions: MVG_PIPE_DEPTH of them, a third of the context each (fixed, so that submissions with different
     * chunk sizes agree on what a region is); a chunk is at most a region and at most an eighth of the batch, so
     * that H2D, kernels and D2H of neighbouring chunks overlap even when the whole batch would fit in one region. */
    int depth = MVG_PIPE_DEPTH;
    if (ctx->max_pics < depth) depth = 1;
    const int region = std::max(1, ctx->max_pics / depth);
    int chunk = std::min(region, std::max(1, (n_pics + 7) / 8));
    /* thumbnails: little to copy back, and a launch over few pictures is bound by the critical path through a picture
     * (about 1 ms whatever their number): fewer, larger chunks */
    if (scale > 1 && !yuv_out) chunk = std::min(region, std::max(chunk, std::min(n_pics, 128)));
    if (ctx->pipe_chunk > 0) chunk = std::min(ctx->pipe_chunk, region);
    for (int done = 0; done < n_pics; done += chunk, ctx->pipe_idx++) {
        const int cnt = std::min(chunk, n_pics - done);
        const int r = (int)(ctx->pipe_idx % depth), slot0 = r * region;
        const bool used = ctx->region_used[r];
        ctx->region_used[r] = true;
        /* inputs of region r were last read by the compute of the chunk that used it before */
        if (used) CK(ctx, cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_comp[r], 0));
        if (upload(done, slot0, cnt) != MVG_SUCCESS) return MVG_FAILURE;
        CK(ctx, cudaEventRecord(ctx->ev_h2d[r], ctx->s_h2d));
        /* outputs of region r were last read by the D2H of the chunk that used it before */
        CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[r], 0));
        if (used) CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[r], 0));
        if (expand(slot0, cnt) != MVG_SUCCESS) return MVG_FAILURE;
        if (launch_stages(ctx, slot0, cnt, scale, yuv_out != nullptr, ctx->stream, false) != MVG_SUCCESS) return MVG_FAILURE;
        if (yuv_out && launch_planar(ctx, slot0, cnt, ctx->stream) != MVG_SUCCESS) return MVG_FAILURE;
        CK(ctx, cudaEventRecord(ctx->ev_comp[r], ctx->stream));
        CK(ctx, cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_comp[r], 0));
        if (yuv_out)
            CK(ctx, cudaMemcpyAsync(yuv_out + (size_t)done * yuv_sz, ctx->d_yuv + (size_t)slot0 * yuv_sz,
                                    (size_t)cnt * yuv_sz, cudaMemcpyDeviceToHost, ctx->s_d2h));
        if (rgb_out)
            CK(ctx, cudaMemcpyAsync(rgb_out + (size_t)done * rgb_sz, ctx->d_rgb + (size_t)slot0 * rgb_sz,
                                    (size_t)cnt * rgb_sz, cudaMemcpyDeviceToHost, ctx->s_d2h));
        CK(ctx, cudaEventRecord(ctx->ev_d2h[r], ctx->s_d2h));
    }
    CK(ctx, cudaEventRecord(done_ev, ctx->s_d2h));
    return MVG_SUCCESS;
}

extern "C" int mvg_wait(mvg_ctx *ctx, mvg_ticket ticket)
{
    MvgRange nvtx_range("mvg_wait");
    if (!ctx) return MVG_FAILURE;
    if (ticket < 0 || ticket >= MVG_MAX_TICKETS || !ctx->tickets[ticket].busy) return fail(ctx, "mvg_wait: ticket %d is not in flight", (int)ticket);
    CK(ctx, cudaSetDevice(ctx->device));
    ctx->tickets[ticket].busy = false;
    CK(ctx, cudaEventSynchronize(ctx->tickets[ticket].done));
    return MVG_SUCCESS;
}

extern "C" int mvg_poll(mvg_ctx *ctx, mvg_ticket ticket, int *done)
{
    if (!ctx || !done) return MVG_FAILURE;
    if (ticket < 0 || ticket >= MVG_MAX_TICKETS || !ctx->tickets[ticket].busy) return fail(ctx, "mvg_poll: ticket %d is not in flight", (int)ticket);
    const cudaError_t e = cudaEventQuery(ctx->tickets[ticket].done);
    if (e != cudaSuccess && e != cudaErrorNotReady) return fail(ctx, "mvg_poll: %s", cudaGetErrorString(e));
    *done = e == cudaSuccess;
    return MVG_SUCCESS;
}

extern "C" int mvg_submit(mvg_ctx *ctx, const mvg_batch *b, uint8_t *yuv_out, uint8_t *rgb_out, int rgb_scale, mvg_ticket *ticket)
{
    MvgRange nvtx_range("mvg_submit");
    if (!ctx) return MVG_FAILURE;
    if (!ticket) return fail(ctx, "mvg_submit: ticket is NULL");
    if (check_batch(ctx, b, "mvg_submit") != MVG_SUCCESS) return MVG_FAILURE;
    if (!ctx->have_sps) return fail(ctx, "mvg_submit: mvg_set_sps() has not been called");
    if (b->n_pics < 1) return fail(ctx, "mvg_submit: empty batch");
    if (rgb_out && rgb_scale < 1) return fail(ctx, "mvg_submit: rgb_out given but rgb_scale < 1");
    CK(ctx, cudaSetDevice(ctx->device));
    if (take_ticket(ctx, ticket) != MVG_SUCCESS) return MVG_FAILURE;
    MvgTicket &t = ctx->tickets[*ticket];
    const mvg_batch batch = *b;     /* the arrays must stay valid until mvg_wait(); the structure need not */
    if (decode_pipeline(ctx, batch.n_pics, yuv_out, rgb_out, rgb_scale, t.done,
                        [&](int done, int slot0, int cnt) { return upload_async(ctx, &batch, done, slot0, cnt, ctx->s_h2d); },
                        [&](int, int) { return (int)MVG_SUCCESS; }) != MVG_SUCCESS) return MVG_FAILURE;
    t.busy = true;
    return MVG_SUCCESS;
}

extern "C" int mvg_decode_host(mvg_ctx *ctx, const mvg_batch *b, uint8_t *yuv_out, uint8_t *rgb_out, int rgb_scale)
{
    mvg_ticket t;
    if (mvg_submit(ctx, b, yuv_out, rgb_out, rgb_scale, &t) != MVG_SUCCESS) return MVG_FAILURE;
    return mvg_wait(ctx, t);
}

/* ------------------------------------------------------------------------- */
/* packed transfer format                                                      */

extern "C" int mvg_pack_batch(const int16_t *coeff, int n_pics, int n_mbs, uint32_t *nz_blocks, uint32_t *word_off,
                              uint64_t *pic_off, uint16_t *words, size_t words_capacity, int n_threads)
{
    MvgRange nvtx_range("mvg_pack_batch");
    if (!coeff || !nz_blocks || !word_off || !pic_off || !words || n_pics < 1 || n_mbs < 1) return MVG_FAILURE;
    try {               /* allocation or thread creation may throw: nothing may unwind through the C ABI */
    /* pass 1 (threaded over pictures): chunk bitmaps and per-macroblock sizes -> offsets inside the picture */
    std::vector<uint64_t> pic_words((size_t)n_pics);
    auto pass1 = [&](int p) {
        uint64_t off = 0;
        for (int m = 0; m < n_mbs; m++) {
            const size_t mb = (size_t)p * n_mbs + m;
            const int16_t *c = coeff + mb * 384;
            uint32_t nzb = 0;
            unsigned words_here = 0;
            for (int b = 0; b < 24; b++) {
                unsigned cnt = 0;
                for (int k = 0; k < 16; k++) cnt += c[b * 16 + k] != 0;
                if (cnt) { nzb |= 1u << b; words_here += 1 + cnt; }
            }
            nz_blocks[mb] = nzb;
            word_off[mb] = (uint32_t)off;
            off += words_here;
        }
        pic_words[(size_t)p] = off;
    };
    auto pass2 = [&](int p) {
        for (int m = 0; m < n_mbs; m++) {
            const size_t mb = (size_t)p * n_mbs + m;
            const int16_t *c = coeff + mb * 384;
            const uint32_t nzb = nz_blocks[mb];
            uint16_t *w = words + pic_off[p] + word_off[mb];
            uint16_t *lv = w + __builtin_popcount(nzb);
            for (int b = 0; b < 24; b++) {
                if (!((nzb >> b) & 1u)) continue;
                unsigned mask = 0;
                for (int k = 0; k < 16; k++)
                    if (c[b * 16 + k]) { mask |= 1u << k; *lv++ = (uint16_t)c[b * 16 + k]; }
                *w++ = (uint16_t)mask;
            }
        }
    };
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, n_pics));
    auto run = [&](auto &fn) {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++)
            th.emplace_back([&, t] { for (int p = t; p < n_pics; p += nt) fn(p); });
        for (auto &x : th) x.join();
    };
    run(pass1);
    pic_off[0] = 0;
    for (int p = 0; p < n_pics; p++) pic_off[p + 1] = pic_off[p] + pic_words[(size_t)p];
    if (pic_off[n_pics] > words_capacity) return MVG_FAILURE;
    run(pass2);
    } catch (...) {
        return MVG_FAILURE;
    }
    return MVG_SUCCESS;
}

extern "C" int mvg_submit_packed(mvg_ctx *ctx, const mvg_packed_batch *b, uint8_t *yuv_out, uint8_t *rgb_out, int rgb_scale, mvg_ticket *ticket)
{
    MvgRange nvtx_range("mvg_submit_packed");
    if (!ctx) return MVG_FAILURE;
    if (!ticket) return fail(ctx, "mvg_submit_packed: ticket is NULL");
    if (!b) return fail(ctx, "mvg_decode_host_packed: batch is NULL");
    if (!b->mb_kind || !b->i16_mode || !b->chroma_mode || !b->qp_y || !b->luma_modes || !b->nz_blocks || !b->word_off ||
        !b->pic_off || !b->words)
        return fail(ctx, "mvg_decode_host_packed: a required pointer is NULL");
    if (!ctx->have_sps) return fail(ctx, "mvg_decode_host_packed: mvg_set_sps() has not been called");
    if (b->n_pics < 1) return fail(ctx, "mvg_decode_host_packed: empty batch");
    if (rgb_out && rgb_scale < 1) return fail(ctx, "mvg_decode_host_packed: rgb_out given but rgb_scale < 1");
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t n = ctx->n_mb();
    if (!ctx->d_words) {
        const size_t cap = ctx->n_mb_max() * (size_t)ctx->max_pics;
        CK(ctx, dalloc(ctx, &ctx->d_nzb, cap, "nz_blocks"));
        CK(ctx, dalloc(ctx, &ctx->d_woff, cap, "word_off"));
        CK(ctx, dalloc(ctx, &ctx->d_words, cap * MVG_PACKED_WORDS_PER_MB, "words"));
        CK(ctx, dalloc(ctx, &ctx->d_picbase, (size_t)ctx->max_pics * 2 + 2, "pic_base"));
    }
    if (take_ticket(ctx, ticket) != MVG_SUCCESS) return MVG_FAILURE;
    MvgTicket &tk = ctx->tickets[*ticket];
    if (b->n_pics > tk.cap) {
        if (tk.h_picbase) cudaFreeHost(tk.h_picbase);
        tk.h_picbase = nullptr; tk.cap = 0;
        CK(ctx, cudaHostAlloc((void **)&tk.h_picbase, ((size_t)b->n_pics * 2 + 2) * sizeof(uint64_t), cudaHostAllocDefault));
        tk.cap = b->n_pics;
    }
    for (int p = 0; p < b->n_pics; p++)
        if (b->pic_off[p + 1] < b->pic_off[p] || b->pic_off[p + 1] - b->pic_off[p] > (uint64_t)n * MVG_PACKED_WORDS_PER_MB)
            return fail(ctx, "mvg_decode_host_packed: picture %d has an impossible word count", p);
    auto upload = [&](int done, int slot0, int cnt) -> int {
        const size_t o = (size_t)slot0 * n, s = (size_t)done * n, c = (size_t)cnt * n;
        cudaStream_t st = ctx->s_h2d;
        CK(ctx, cudaMemcpyAsync(ctx->d_kind + o, b->mb_kind + s, c, cudaMemcpyHostToDevice, st));
        CK(ctx, cudaMemcpyAsync(ctx->d_i16 + o, b->i16_mode + s, c, cudaMemcpyHostToDevice, st));
        CK(ctx, cudaMemcpyAsync(ctx->d_cm + o, b->chroma_mode + s, c, cudaMemcpyHostToDevice, st));
        CK(ctx, cudaMemcpyAsync(ctx->d_qp + o, b->qp_y + s, c, cudaMemcpyHostToDevice, st));
        CK(ctx, cudaMemcpyAsync(ctx->d_modes + o * 16, b->luma_modes + s * 16, c * 16, cudaMemcpyHostToDevice, st));
        CK(ctx, cudaMemcpyAsync(ctx->d_nzb + o, b->nz_blocks + s, c * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CK(ctx, cudaMemcpyAsync(ctx->d_woff + o, b->word_off + s, c * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        const uint64_t w0 = b->pic_off[done], w1 = b->pic_off[done + cnt], base = (uint64_t)o * MVG_PACKED_WORDS_PER_MB;
        if (w1 > w0)
            CK(ctx, cudaMemcpyAsync(ctx->d_words + base, b->words + w0, (size_t)(w1 - w0) * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
        /* cnt + 1 offsets per chunk (the last one is the end of its words), at twice the picture index so that
         * neighbouring chunks / slot regions do not overlap; host entries are not reused before the final sync */
        uint64_t *hb = tk.h_picbase + 2 * (size_t)done;
        for (int i = 0; i <= cnt; i++) hb[i] = base + (b->pic_off[done + i] - w0);
        CK(ctx, cudaMemcpyAsync(ctx->d_picbase + 2 * (size_t)slot0, hb, (size_t)(cnt + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        return MVG_SUCCESS;
    };
    auto expand = [&](int slot0, int cnt) -> int {
        K0Params p;
        const size_t o = (size_t)slot0 * n;
        p.nz_blocks = ctx->d_nzb + o; p.word_off = ctx->d_woff + o; p.pic_base = ctx->d_picbase + 2 * (size_t)slot0;
        p.words = ctx->d_words; p.coeff = ctx->d_coeff + o * 384; p.n_mbs = (long long)cnt * (long long)n; p.mbs_per_pic = (int)n;
        const long long warps = p.n_mbs;
        const int grid = (int)std::min<long long>((warps + 7) / 8, (long long)ctx->sm_count * 16);
        k0_expand_levels<<<grid, 256, 0, ctx->stream>>>(p);
        CK(ctx, cudaGetLastError());
        return MVG_SUCCESS;
    };
    if (decode_pipeline(ctx, b->n_pics, yuv_out, rgb_out, rgb_scale, tk.done, upload, expand) != MVG_SUCCESS) return MVG_FAILURE;
    tk.busy = true;
    return MVG_SUCCESS;
}

extern "C" int mvg_decode_host_packed(mvg_ctx *ctx, const mvg_packed_batch *b, uint8_t *yuv_out, uint8_t *rgb_out, int rgb_scale)
{
    mvg_ticket t;
    if (mvg_submit_packed(ctx, b, yuv_out, rgb_out, rgb_scale, &t) != MVG_SUCCESS) return MVG_FAILURE;
    return mvg_wait(ctx, t);
}

