/*
 * mvg_fused.cuh -- the fused reconstruction kernel: levels in, reconstructed picture out, one launch.
 *
 *   kf_recon<KF_OUT_TILES>  dequantisation + inverse transforms + intra prediction + residual add
 *                           -> macroblock tiles (what kernel 3 / kernel 4 read)
 *   kf_recon<KF_OUT_RGB>    the same + 4:2:0 -> RGB24 of mb_to_rgb() (export_utils.c:266-303)
 *                           -> the RGB picture; nothing else is written
 *   kf_recon<KF_OUT_RGBS>   the same + box downscale by 2, 4, 8 or 16 -> the RGB thumbnail; nothing else is written
 *
 * This is kernel 1 folded into the row loop of kernel 2 (and kernel 3 folded into its write-out), the way the
 * reference itself interleaves them per block (h264_intra_prediction.c:173,954,1929,2319 call
 * transform4x4_luma / transform8x8_luma / transform16x16_luma / transform4x4_chroma of h264_transform.c:121-402
 * right after predicting).  The residual never exists in HBM: per picture the kernel reads the 789 B of
 * structure-of-arrays per macroblock and writes 384 B (tiles) or 768 B (RGB24), against 31.8 MB for the three
 * separate kernels.  All device code it is made of lives in mvg_kernels.cuh and is shared with kernels 1-3
 * (mvg_xf_group, k2_luma4/8/16, k2_chroma), so the split pipeline and the fused one cannot drift apart.
 */
#pragma once

#include <cstddef>

#include "mvg_kernels.cuh"

#ifndef KF_GROUP
#define KF_GROUP 4          /* macroblocks transformed together: the compaction of non-zero blocks over a group is what
                               makes the transform stage cheap (2 macroblocks: 6.1 ms, 4: 5.4 ms per 1000 pictures) */
#endif
#ifndef KF_WARPS
#define KF_WARPS 25         /* warps per CTA (one CTA per SM), RGB mode: 8 KB of shared memory per warp (the group's RGB24 staging is 3.3 KB of it) */
#endif
#ifndef KF_WARPS_TILES
#define KF_WARPS_TILES 28   /* tiles mode: 4.75 ms per 1000 pictures at 28 warps (24: 4.84, 30: 4.82) */
#endif
#ifndef KF_WARPS_RGBS
#define KF_WARPS_RGBS 28    /* thumbnail mode */
#endif
#define KF_WARPS_OF(out) ((out) == 1 ? KF_WARPS : (out) == 2 ? KF_WARPS_RGBS : KF_WARPS_TILES)
#ifndef KF_COMPACT8
#define KF_COMPACT8 false   /* true: Intra8x8 blocks through one run-time-indexed copy of the code (smaller, 4 % slower) */
#endif
#ifndef KF_POLL_NS
#define KF_POLL_NS 1000     /* sleep between two looks at the row above while waiting for the distance */
#endif
#ifndef KF_STAGGER
#define KF_STAGGER 12       /* distance (macroblocks) a row keeps from the row above in a launch that has more rows than the
                               GPU has warps (KFParams::stagger; smaller launches follow at the minimum): waited for at the row start and
                               again whenever the row catches up; the halo prefetch reaches 10 macroblocks ahead */
#endif
#ifndef KF_MBS
#define KF_MBS 392          /* int16 per macroblock slot of the level buffer: 768 bytes of levels + 16 of padding, so that the DC
                               levels of the group's macroblocks fall into different banks (see mvg_xf_group) */
#endif
#define KF_OUT_TILES 0
#define KF_OUT_RGB   1
#define KF_OUT_RGBS  2      /* RGB24 thumbnails: the s x s box average of the RGB picture, s = 2, 4, 8, 16 */
#define KF_RGBS_PITCH 96     /* staging row of a group in thumbnail mode: at most 4 macroblocks x 8 pixels x 3 bytes (s = 2) */
/* RGB staging of a PAIR of macroblocks (only with KF_WO_GROUP = 0): 16 rows of 2 x 48 bytes + 16 of padding */
#define KF_RGB_STRIDE 112
/* KF_WO_GROUP: the RGB24 of a whole group of four macroblocks is staged and written out once per group.  Staging: 16
 * rows of 4 x 48 bytes + 16 of padding = 13 pieces of 16 bytes (odd: eight consecutive rows start in eight different
 * 16-byte bank groups), even picture rows in staging rows 0..7, odd ones in 8..15.  The conversion's lane 4 q + h stores
 * words 3 h + w of picture rows 2 q, 2 q + 1 = staging rows q, q + 8: banks 52 q + 3 h + w, all 32 different.  The
 * write-out moves 64-byte pieces: the eight lanes of a quarter warp -- the unit in which 16-byte accesses are processed,
 * for the global stores too: a quarter that covers eight picture rows costs eight tag look-ups -- take 4 x 16 bytes of
 * staging row s and of s + 4 (13 s and 13 (s + 4) pieces differ by 4 modulo 8: all eight bank groups), i.e. two picture
 * rows, two 128-byte lines.  Every address is a lane constant plus an immediate. */
#ifndef KF_WO_GROUP
#define KF_WO_GROUP (KF_GROUP == 4)
#endif
#define KF_RGB_GSTRIDE 208
#ifndef KF_HALO_GROUP
#define KF_HALO_GROUP (KF_GROUP == 4)   /* the halo words of the row above are validated and parked in shared memory once
                                           per group of four macroblocks (0: checked per macroblock, kept in registers) */
#endif

struct KFParams {
    const uint8_t *mb_kind, *i16_mode, *chroma_mode, *luma_modes;   /* [slot][n_mb](x16), slot 0 */
    const int8_t  *qp_y;
    const int16_t *coeff;       /* [slot][n_mb][384] levels                                           */
    uint8_t       *tiles;       /* [slot][n_mb][384] (KF_OUT_TILES)                                   */
    uint8_t       *rgb;         /* [slot][3 * W * H] (KF_OUT_RGB), [slot][3 * (W / s) * (H / s)] (KF_OUT_RGBS) */
    uint2         *halo;        /* [slot][n_mb][8] bottom sample line of every macroblock + epoch     */
    int           *work;        /* work counter of this launch (starts at 0)                          */
    const MvgTables *tab;
    const MvgLuts   *luts;
    unsigned        epoch;
    int w_mbs, h_mbs, first_slot, n_pics, group;
    int stagger;                /* distance in macroblocks a row keeps from the row above (KF_STAGGER; less for small launches) */
    unsigned        sel[4];
    unsigned long long *stats;  /* DEV (-DKF_STATS): wait accounting, 16 counters */
};

#ifdef KF_STATS
#define KF_STAT(...) __VA_ARGS__
#else
#define KF_STAT(...)
#endif

/* Member order matters (see K2WarpSmem): lanes without a block in an Intra4x4 step read up to 64 bytes below
 * the residual and 140 bytes below lt[] (and past the end of lt[] into ct[]); all of that stays inside this record. */
template <int OUT>
struct KFWarpSmemT {
    union {
        MvgXfScratch<KF_GROUP> x;                       /* transform stage                                      */
        uint8_t rgb[OUT == 0 ? 16 : OUT == 2 ? 8 * KF_RGBS_PITCH : KF_WO_GROUP ? 16 * KF_RGB_GSTRIDE : 16 * KF_RGB_STRIDE];    /* RGB24 rows of a macroblock group / pair (prediction stage, KF_OUT_RGB) */
    } u;
    MVG_CANARY(c0)
    __align__(128) int16_t tile[KF_GROUP * KF_MBS];        /* levels in -> residual in place; slot j is refilled with macroblock j
                                                           of the next group as soon as macroblock j has been predicted */
    __align__(8) uint64_t mbar;
    __align__(16) uint32_t hrow[36];                    /* KF_HALO_GROUP: the bottom sample line of the four macroblocks above this
                                                           group and the first two words of the next one */
    MVG_CANARY(c1)
    __align__(16) uint8_t lt[MVG_LT_ROWS * MVG_LT_STRIDE];
    /* RGB modes: the chroma tiles start an ODD number of words behind the luma tile.  The hand-over of the right column
     * (x = 15 / 7 -> x = -1) is one byte per lane, rows 0..15 of luma in lanes 0..15 and the 2 x 8 chroma rows in lanes
     * 16..31: luma words are 10 l + const apart (16 odd banks), chroma words 6 r + const (16 even banks, Cr 80 words behind
     * Cb) -- conflict-free only if the two sets have different parity.  The tiles mode reads chroma rows as 8-byte pieces
     * and keeps the aligned layout. */
    uint8_t ct_pad[4];
    alignas(OUT == 0 ? 16 : 4) uint8_t ct[2][MVG_CT_PLANE];
    MVG_CANARY(c2)
    __align__(16) uint8_t n8[MVG_N8_BYTES];
    MVG_CANARY(c3)
#ifdef KF_REC_PAD        /* DEV: a larger record without any use of it (shared memory carve-out experiments) */
    uint8_t pad[KF_REC_PAD];
#endif
};

#define KF_LUT_BYTES  ((sizeof(MvgLuts) + 127) / 128 * 128)
#define KF_TAB_BYTES  ((sizeof(MvgXfTables) + 127) / 128 * 128)
#define KF_SMEM_BYTES(out) (sizeof(KFWarpSmemT<out>) * KF_WARPS_OF(out) + 2048 + KF_LUT_BYTES + KF_TAB_BYTES)

/* Persistent warps, one CTA per SM.  A work item is one macroblock row of one picture (claimed from an atomic
 * counter, rows of a picture in order, pictures interleaved: see k2_wavefront for the dependency protocol, which
 * is unchanged: flag-in-data bottom lines, no fences).  A warp walks its row in groups of KF_GROUP macroblocks:
 *   1. the group's levels arrive by bulk asynchronous copies (TMA 1-D + mbarrier), 768 bytes per macroblock, each
 *      requested as soon as its slot is free: right after the macroblock that held it a group earlier is predicted;
 *   2. mvg_xf_group() turns them into the residual in place (kernel 1's code);
 *   3. the bottom sample line of the four macroblocks above (+ the first two words of the next one), requested a group
 *      ago with one 16-byte relaxed load per lane, is validated once and parked in shared memory;
 *   4. each macroblock is predicted and reconstructed in the warp's tile (kernel 2's code), its bottom line
 *      published, and then either stored as a 384-byte tile or converted to RGB24 into the group's staging area;
 *   5. RGB mode: after the group's last macroblock its 16 rows x 192 bytes leave as 16-byte stores, 64 bytes of two
 *      picture rows per quarter warp.
 * What the kernel runs out of first is the shared-memory/LSU data pipe and instruction issue, both at about 80 % (ncu,
 * profiles/): hence conflict-free layouts everywhere, and nothing in the row loop that the compiler could turn into a
 * special-register read (the opaque base addresses below).  The body is 49 KB of code, just under what the instruction cache holds for
 * 25 warps at different places of it: unrolling one 95-instruction loop three times costs 17 % (profiles/r02_notes.md). */
#ifdef KF_MAXREG
#define KF_BOUNDS __maxnreg__(KF_MAXREG)        /* explicit register budget (the block size is given at launch) */
#else
#define KF_BOUNDS __launch_bounds__(KF_WARPS_OF(OUT) * 32, 1)
#endif
template <int OUT, int SL = 2>      /* SL: log2 of the thumbnail scale (KF_OUT_RGBS only).  A template parameter, not a kernel
                                       argument: with all four scales in one body the code outgrows the instruction cache */
__global__ void KF_BOUNDS
kf_recon(KFParams p)
{
    extern __shared__ __align__(128) uint8_t kf_smem[];
    typedef KFWarpSmemT<OUT> KFWarpSmem;        /* the tiles mode has no RGB staging area */
    const int lane = mvg_lane();
    const unsigned wid = __shfl_sync(MVG_FULL, threadIdx.x >> 5, 0);       /* warp-uniform by construction */
    /* layout: the tap tables sit on the first 2 KB boundary (so that (mode << 7) can be OR-ed into a lane's
     * table address), the dequantisation tables behind them; warp records fill the space before, the rest follow */
    const unsigned base = mvg_smem_u32(kf_smem);
    /* The addresses everything hangs on -- the tables and this warp's record -- as opaque warp-uniform 32-bit shared
     * addresses.  Left to itself the compiler carries them as (shared window base) + (offset) in two uniform registers,
     * spends two instructions per lane-dependent address on adding them up and, worse, re-derives the window base inside
     * the row loop from a special register (S2R SR_CgaCtaId: tens of cycles each time). */
    const unsigned lut_addr = __shfl_sync(MVG_FULL, mvg_keep((base + 2047u) & ~2047u), 0);
    const unsigned n_before = (lut_addr - base) / (unsigned)sizeof(KFWarpSmem);
    MvgLuts *luts = reinterpret_cast<MvgLuts *>(__cvta_shared_to_generic(lut_addr));
    MvgXfTables &T = *reinterpret_cast<MvgXfTables *>(__cvta_shared_to_generic(lut_addr + (unsigned)KF_LUT_BYTES));
    unsigned rec_addr = base + (wid < n_before ? wid * (unsigned)sizeof(KFWarpSmem)
                                               : (lut_addr - base) + (unsigned)(KF_LUT_BYTES + KF_TAB_BYTES) + (wid - n_before) * (unsigned)sizeof(KFWarpSmem));
    rec_addr = __shfl_sync(MVG_FULL, mvg_keep(rec_addr), 0);
    KFWarpSmem &s = *reinterpret_cast<KFWarpSmem *>(__cvta_shared_to_generic(rec_addr));
    for (int i = threadIdx.x; i < (int)(KF_LUT_BYTES / 16); i += blockDim.x)
        reinterpret_cast<uint4 *>(luts)[i] = __ldg(reinterpret_cast<const uint4 *>(p.luts) + i);
    mvg_xf_load_tables(T, p.tab);
    if (lane == 0) mvg_mbar_init(&s.mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const int W = p.w_mbs, H = p.h_mbs, n_mb = W * H;
    const int total = p.n_pics * H;
    const unsigned epoch = p.epoch;
    const int n_groups = (W + KF_GROUP - 1) / KF_GROUP;

    /* per-lane constants (as in k2_wavefront) ------------------------------------------ */
    K2Ctx c;
#ifdef MVG_CHECKED
    c.rec_lo = reinterpret_cast<const uint8_t *>(&s); c.rec_hi = c.rec_lo + sizeof(KFWarpSmem);
    if (lane < 4) s.c0[lane] = s.c1[lane] = s.c2[lane] = s.c3[lane] = MVG_CANARY_WORD;
    __syncwarp();
#endif
    c.lt = s.lt; c.ct = &s.ct[0][0]; c.n8 = s.n8; c.lut8 = reinterpret_cast<const uint8_t *>(&luts->lut8[0][lane]);
    c.lane = lane;
    c.sel = p.sel;
    c.resid = reinterpret_cast<const uint8_t *>(s.tile);
    {
        const int half = lane >> 4, pix = lane & 15, px = pix & 3, py = pix >> 2;
        c.lut4 = lut_addr + (unsigned)lane * 4u;
        c.h4 = half ? 0u : (unsigned)(4 * MVG_LT_STRIDE - 8);
        c.s4 = py * MVG_LT_STRIDE + px + (int)c.h4;
        c.r4odd = pix * 2 + (half ? 64 : 0);
        c.r4even = pix * 2 + (half ? -64 : 0);
        c.m4c = half ? 0x10100010u : 0x10001000u;
        c.m4b = half ? 0x00000100u : 0x00000011u;
        c.m4cc = half ? 0x00001000u : 0u;
        const int n = lane < 25 ? lane : 24;
        if (n < 8)       c.n8tr = (7 - n) * MVG_LT_STRIDE - 1;
        else if (n == 8) c.n8tr = -MVG_LT_STRIDE - 1;
        else             c.n8tr = -MVG_LT_STRIDE + (n - 9);
        c.n8notr = n > 16 ? -MVG_LT_STRIDE + 7 : c.n8tr;
        const int x8 = ((lane >> 3) & 1) * 4 + (lane & 1) * 2, y8 = (lane >> 4) * 4 + ((lane >> 1) & 3);
        c.s8 = y8 * MVG_LT_STRIDE + x8;
        c.fixA = lane == 7 ? 0x10u : lane == 8 ? 0x22u : lane == 9 ? 0x20u : 0u;
        c.fixB = lane == 7 ? 0x04u : lane == 8 ? 0x05u : lane == 9 ? 0x08u : 0u;
        c.fixD = lane == 7 ? 0x01u : lane == 9 ? 0x02u : 0u;
        /* the Intra4x4 step constants stay in registers (mvg_keep): without this the compiler re-derives them from the
         * lane index in every one of the ten steps.  6.58 -> 6.41 ms per 1000 pictures; keeping the Intra8x8 and the RGB
         * output offsets as well costs more in register pressure than it saves (6.51 / 7.11 ms) */
        c.lut4 = mvg_keep(c.lut4); c.h4 = mvg_keep(c.h4); c.s4 = mvg_keep(c.s4); c.r4odd = mvg_keep(c.r4odd); c.r4even = mvg_keep(c.r4even);
    }
    /* sample row -1 of the tiles comes from the halo words of the row above: lanes 0..3 luma x = 4*lane,
     * 4,5 Cb, 6,7 Cr of the macroblock above, lanes 8,9 luma x = 16..23 of the macroblock above-right */
    uint8_t *const halo_top = lane < 4 ? s.lt + K2_TO(lane * 4, -1)
                            : lane < 8 ? s.ct[(lane >> 1) & 1] + K2_CO((lane & 1) * 4, -1)
                                       : s.lt + K2_TO(16 + (lane & 1) * 4, -1);
    const uint8_t *const halo_bot = lane < 4 ? halo_top + 16 * MVG_LT_STRIDE : halo_top + 8 * MVG_CT_STRIDE;
    /* row -1, x = 15 (luma) / 7 (chroma) -> x = -1 of the next macroblock: lanes 3, 5, 7 hold the last word of the
     * luma / Cb / Cr line above; the other lanes copy an unused byte onto itself so that the move needs no predicate */
    const uint8_t *const cn_src = (lane == 3 || lane == 5 || lane == 7) ? halo_top + 3 : s.lt;
    uint8_t *const cn_dst = lane == 3 ? s.lt + K2_TO(-1, -1) : lane == 5 ? s.ct[0] + K2_CO(-1, -1)
                          : lane == 7 ? s.ct[1] + K2_CO(-1, -1) : s.lt;
    /* what a lane reads of the finished macroblock, and where the last byte of it goes as the left neighbour column
     * of the next macroblock (x = 15 -> x = -1 luma, x = 7 -> x = -1 chroma):
     *   tiles: lanes 0..15 one luma row (two 8-byte pieces), 16..23 Cb rows, 24..31 Cr rows (8 bytes)
     *   RGB:   lane = 4 q + h: luma samples 4h..4h+3 of rows 2q and 2q+1, Cb and Cr samples 2h, 2h+1 of row q -- the
     *          four pixels of a row and the row below share their two chroma samples, so the chroma terms of the
     *          conversion are computed once per lane; the h = 3 lanes hold x = 15 of both luma rows and x = 7 of
     *          the chroma row                                                                                      */
    const uint8_t *wo_src, *wo_csrc = nullptr;
    uint8_t *lc_dst, *lc_cdst = nullptr, *rgb_dst = nullptr;
    const uint8_t *lc_src = nullptr;
    int wo_off = 0;
    if (OUT == KF_OUT_TILES) {
        wo_src = lane < 16 ? s.lt + K2_TO(0, lane) : s.ct[(lane >> 3) & 1] + K2_CO(0, lane & 7);
        wo_off = lane < 16 ? lane * 16 : 256 + (lane - 16) * 8;
        lc_dst = lane < 16 ? s.lt + K2_TO(-1, lane) : s.ct[(lane >> 3) & 1] + K2_CO(-1, lane & 7);
    } else {
        const int q = lane >> 2, h = lane & 3;
        wo_src = s.lt + K2_TO(4 * h, 2 * q);
        wo_csrc = s.ct[0] + K2_CO(2 * h, q);
        lc_dst = s.lt + K2_TO(-1, 2 * q);
        lc_cdst = s.ct[0] + K2_CO(-1, q);
        /* the left neighbour column of the next macroblock by a copy of its own, one sample per lane (rows 0..15 of
         * luma, 0..7 of Cb, 0..7 of Cr): one load and one store instead of four predicated byte stores */
        lc_dst = lane < 16 ? s.lt + K2_TO(-1, lane) : s.ct[(lane >> 3) & 1] + K2_CO(-1, lane & 7);
        lc_src = lc_dst + (lane < 16 ? 16 : 8);
        rgb_dst = KF_WO_GROUP ? s.u.rgb + q * KF_RGB_GSTRIDE + 12 * h : s.u.rgb + 2 * q * KF_RGB_STRIDE + 12 * h;
    }
    const unsigned lc_sel = lane < 16 ? 7u : 3u;            /* tiles: byte 3 of the second / first 8-byte piece */
    const int pitch = 48 * W;                               /* bytes per RGB24 picture row */

    MvgSideInfo side;
    side.init(lane, p.mb_kind, p.i16_mode, p.chroma_mode, p.luma_modes, p.qp_y);
    unsigned parity = 0;            /* phase parity of the mbarrier: one phase per group */
    KF_STAT(unsigned long long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};)

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(p.work, 1);
        item = __shfl_sync(MVG_FULL, item, 0);
        if (item >= total) break;
        const int g0 = item / (p.group * H);
        const int gsize = min(p.group, p.n_pics - g0 * p.group);
        const int within = item - g0 * p.group * H;
        const int row = within / gsize;
        const int slot = p.first_slot + g0 * p.group + (within - row * gsize);
        const size_t mb0 = (size_t)slot * n_mb + (size_t)row * W;           /* first macroblock of the row */
        const int16_t *lv_row = p.coeff + mb0 * 384;                        /* levels of this row */

        /* group 0: levels and side information.  The buffer was last touched by this warp's generic-proxy
         * accesses (residual of an earlier group): order them before the asynchronous write. */
        if (lane == 0) {
            const int n0 = min(KF_GROUP, W);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mvg_mbar_expect_tx(&s.mbar, (unsigned)n0 * 768u);
            if (KF_MBS == 384) mvg_bulk_load(s.tile, lv_row, (unsigned)n0 * 768u, &s.mbar);
            else for (int j = 0; j < n0; j++) mvg_bulk_load(s.tile + j * KF_MBS, lv_row + j * 384, 768u, &s.mbar);
        }
        unsigned nmeta = side.load(lane, (long long)mb0, min(KF_GROUP, W));

        /* thumbnail mode: `per` output pixels per macroblock side, picture rows of W * per * 3 bytes */
        constexpr int sl = SL, per = 16 >> sl;
        const int opitch = W * per * 3;
        uint8_t *wo_run = OUT == KF_OUT_TILES ? p.tiles + mb0 * 384 + wo_off
                        : OUT == KF_OUT_RGB   ? p.rgb + (size_t)slot * ((size_t)n_mb * 768) + (size_t)row * 16 * pitch
                                              : p.rgb + ((size_t)slot * H + row) * ((size_t)per * opitch);
        uint2 *hm_run = p.halo + mb0 * 8 + lane;
        const bool availB = row > 0, publish = row < H - 1;
        const int hwords = W * 8;                               /* halo words of a macroblock row */
#if KF_HALO_GROUP
        /* lanes 0..16: words 2 l and 2 l + 1 of the 32 + 2 a group needs; lanes that have nothing to load keep the epoch */
        const uint4 *ha_run = reinterpret_cast<const uint4 *>(p.halo + (mb0 - W) * 8) + lane;
        uint4 hq = make_uint4(0, epoch, 0, epoch);
#else
        const uint2 *ha_run = p.halo + (mb0 - W) * 8 + lane;           /* group of macroblock mx (at hj == 0) */
        uint2 qa = make_uint2(0, epoch), qb = make_uint2(0, epoch);
#endif
        if (availB) {
            if (p.stagger > 0) {
                /* Slack between the rows of a picture: wait here, once, until the row above is KF_STAGGER macroblocks
                 * ahead.  Rows that follow each other at the minimum distance (two macroblocks) run in lock step and
                 * every burst of the row above -- it transforms KF_GROUP macroblocks, then predicts them -- stalls all
                 * rows below in turn; with slack the per-macroblock check further down almost never fails. */
                const uint2 *probe = p.halo + (mb0 - W + min(p.stagger, W - 1)) * 8 + 7;
                KF_STAT(const long long ts = clock64();)
                while (mvg_ld_relaxed_u64(probe).y != epoch) { __nanosleep(2000); KF_STAT(st[0]++;) }
                KF_STAT(st[1] += clock64() - ts;)
            }
#if KF_HALO_GROUP
            if (lane < 17 && 2 * lane < hwords) hq = mvg_ld_relaxed_2u64(ha_run);
#else
            if (lane < hwords) qb = mvg_ld_relaxed_u64(ha_run);     /* becomes qa at macroblock 0 */
#endif
        }
#if !KF_HALO_GROUP
        unsigned okA = 0;
#endif

        KF_STAT(st[5]++; const long long trow = clock64();)
        for (int g = 0; g < n_groups; g++) {
            const unsigned meta = nmeta;
            const int nmb = min(KF_GROUP, W - g * KF_GROUP);
            const int n_next = min(KF_GROUP, W - (g + 1) * KF_GROUP);       /* macroblocks of the next group (<= 0: none) */
            if (n_next > 0) nmeta = side.load(lane, (long long)mb0 + (g + 1) * KF_GROUP, n_next);
            int16_t *tile = s.tile;
            mvg_mbar_wait(&s.mbar, parity);
            parity ^= 1u;
            /* the next phase collects the next group's copies, which are issued one by one below */
            if (n_next > 0 && lane == 0) mvg_mbar_expect_tx(&s.mbar, (unsigned)n_next * 768u);

            /* ---- levels -> residual, in place (kernel 1's stage) ---- */
            mvg_xf_group<KF_GROUP, KF_MBS>(tile, s.u.x, T, meta, nmb, lane);
            const uint4 rec = mvg_ctl_from_meta(meta);      /* valid in lane 8 j */

#if KF_HALO_GROUP
            /* The bottom sample line of the row above for this whole group: 32 words of the four macroblocks above and the
             * first two of the next one (up-right neighbour of the group's last macroblock), requested a group ago with one
             * 16-byte load per lane.  Validated once, parked in shared memory, and the next group's request goes out.  A group
             * therefore starts only when the row above is five macroblocks past the group's first one (the per-macroblock
             * rule is two): with the distance KF_STAGGER that rows keep anyway this costs nothing. */
            if (availB) {
                for (;;) {
                    if (__all_sync(MVG_FULL, hq.y == epoch && hq.w == epoch)) break;
                    /* This row has caught up with the row above.  Do not follow it at the minimum distance: fall back until
                     * the row above is KF_STAGGER macroblocks ahead again, then reload (every word validates itself: the
                     * probe word says nothing about its neighbours). */
                    const uint2 *probe = p.halo + (mb0 - W + min(g * KF_GROUP + max(p.stagger, KF_GROUP + 1), W - 1)) * 8 + 7;
                    KF_STAT(const long long ts = clock64(); st[2]++;)
                    while (mvg_ld_relaxed_u64(probe).y != epoch) { __nanosleep(KF_POLL_NS); KF_STAT(st[3]++;) }
                    KF_STAT(st[4] += clock64() - ts;)
                    if (lane < 17 && g * 32 + 2 * lane < hwords) hq = mvg_ld_relaxed_2u64(ha_run);
                }
                if (lane < 17) *reinterpret_cast<uint2 *>(s.hrow + 2 * lane) = make_uint2(hq.x, hq.z);
                ha_run += 16;
                hq = make_uint4(0, epoch, 0, epoch);
                if (n_next > 0 && lane < 17 && (g + 1) * 32 + 2 * lane < hwords) hq = mvg_ld_relaxed_2u64(ha_run);
                __syncwarp();
            }
#endif

            for (int j = 0; j < nmb; j++) {
                const int mx = g * KF_GROUP + j;
#if !KF_HALO_GROUP
                const int hj = mx & 3;
#endif
                const bool availA = mx > 0, availC = availB && mx < W - 1, availD = availA && availB;

#if KF_HALO_GROUP
                /* sample row -1 of the tiles: lanes 0..7 the macroblock above, lanes 8,9 x = 16..23 (above-right) */
                if (availB && lane < 10) *reinterpret_cast<unsigned *>(halo_top) = s.hrow[8 * j + lane];
                __syncwarp();
#else
                if (availB) {
                    if (hj == 0) {              /* group of four macroblocks above: requested a group ago */
                        qa = qb;
                        okA = __ballot_sync(MVG_FULL, qa.y == epoch);
                    }
                    /* words needed now: the 8 of the macroblock above and, for the up-right neighbour, the first
                     * two of the next one, which sit in qb when this is the last macroblock of the group */
                    /* the common case first: all 32 words of the group are this launch's (lanes past the end of the row
                     * hold the epoch from their initialisation), and so are the two of the next group where they matter */
                    bool all_ok = okA == MVG_FULL;
                    if (hj == 3 && availC) all_ok = all_ok && (__ballot_sync(MVG_FULL, qb.y == epoch) & 3u) == 3u;
                    if (!all_ok) {
                      const unsigned need = availC ? 0x3FFu : 0xFFu;
                      unsigned have = __funnelshift_r(okA, hj == 3 ? __ballot_sync(MVG_FULL, qb.y == epoch) : 0u, 8 * hj);
                      if ((have & need) != need) {
                        /* This row has caught up with the row above.  Do not follow it at the minimum distance: rows in
                         * lock step find the words they prefetch a group ahead stale every time and pay a round trip to
                         * L2 per macroblock.  Fall back until the row above is KF_STAGGER macroblocks ahead again, then
                         * reload; after that the prefetches hit for the next KF_STAGGER - 8 macroblocks at least. */
                        const uint2 *probe = p.halo + (mb0 - W + min(mx + max(p.stagger, 2), W - 1)) * 8 + 7;
                        KF_STAT(const long long ts = clock64(); st[2]++;)
                        while (mvg_ld_relaxed_u64(probe).y != epoch) { __nanosleep(KF_POLL_NS); KF_STAT(st[3]++;) }
                        KF_STAT(st[4] += clock64() - ts;)
                        do {    /* every word validates itself: the probe word says nothing about its neighbours */
                            if ((mx & ~3) * 8 + lane < hwords) qa = mvg_ld_relaxed_u64(ha_run);
                            if ((mx & ~3) * 8 + 32 + lane < hwords) qb = mvg_ld_relaxed_u64(ha_run + 32);
                            okA = __ballot_sync(MVG_FULL, qa.y == epoch);
                            have = __funnelshift_r(okA, __ballot_sync(MVG_FULL, qb.y == epoch), 8 * hj);
                        } while ((have & need) != need);
                      }
                    }
                    /* sample row -1 of the tiles: lanes 0..7 the macroblock above, lanes 8,9 x = 16..23 */
                    const unsigned src = (hj == 3 && lane < 2) ? qb.x : qa.x;
                    const unsigned v = __shfl_sync(MVG_FULL, src, (8 * hj + lane) & 31);
                    if (lane < 10) *reinterpret_cast<unsigned *>(halo_top) = v;
                    if (hj == 3) ha_run += 32;
                }
                __syncwarp();
                if (availB && hj == 0) {
                    /* request the next group of the row above behind everything that reads qb (see k2_wavefront) */
                    qb = make_uint2(0, epoch);
                    if (mx * 8 + 32 + lane < hwords) qb = mvg_ld_relaxed_u64(ha_run + 32);
                }
#endif
                c.resid = reinterpret_cast<const uint8_t *>(tile + j * KF_MBS);
                const unsigned cx = __shfl_sync(MVG_FULL, rec.x, 8 * j), cy = __shfl_sync(MVG_FULL, rec.y, 8 * j);
                const unsigned cz = __shfl_sync(MVG_FULL, rec.z, 8 * j);

                const int kind = cx & 255, i16 = (cx >> 8) & 255, cmode = (cx >> 16) & 255;
                if (kind == MVG_MB_I16x16)    k2_luma16(c, i16, availA, availB);
                else if (kind == MVG_MB_I4x4) k2_luma4(c, cy, cz, availA, availB, availC);
                else                          k2_luma8<KF_COMPACT8>(c, cy, availA, availB, availC, availD);
                k2_chroma(c, cmode, availA, availB);
                __syncwarp();

                /* publish the bottom sample line for the row below */
                if (publish && lane < 8)
                    mvg_st_relaxed_u64(hm_run, *reinterpret_cast<const unsigned *>(halo_bot), epoch);
                hm_run += 8;
                if (OUT == KF_OUT_TILES) {
                    /* the macroblock as one 384-byte tile (coalesced) */
                    const uint2 wa = *reinterpret_cast<const uint2 *>(wo_src);
                    uint2 wb = make_uint2(0u, 0u);
                    if (lane < 16) {
                        wb = *reinterpret_cast<const uint2 *>(wo_src + 8);
                        *reinterpret_cast<uint4 *>(wo_run) = make_uint4(wa.x, wa.y, wb.x, wb.y);
                    } else *reinterpret_cast<uint2 *>(wo_run) = wa;
                    wo_run += 384;
                    *lc_dst = (uint8_t)__byte_perm(wa.y, wb.y, lc_sel);
                } else {
                    /* RGB24 of my 2 x 4 pixels (export_utils.c:300-302 on int16 pairs, see k3_rgb_full) into the staging rows */
                    const unsigned y0 = *reinterpret_cast<const unsigned *>(wo_src);
                    const unsigned y1 = *reinterpret_cast<const unsigned *>(wo_src + MVG_LT_STRIDE);
                    const unsigned cb2 = mvg_pair_lo(*reinterpret_cast<const uint16_t *>(wo_csrc));
                    const unsigned cr2 = mvg_pair_lo(*reinterpret_cast<const uint16_t *>(wo_csrc + MVG_CT_PLANE));
                    *lc_dst = *lc_src;                                  /* x = 15 / 7 -> x = -1 of my row */
                    /* the terms that do not depend on Y: pixels x and x + 2 of a row use chroma samples c and c + 1 */
                    const unsigned rC = __vsub2(((cr2 * 204u) >> 7) & 0x01ff01ffu, 0x00de00deu);                   /* - 222 */
                    const unsigned bC = __vsub2(((cb2 * 129u) >> 6) & 0x03ff03ffu, 0x01140114u);                   /* - 276 */
                    const unsigned gC = __vsub2(__vsub2(0x00870087u, ((cb2 * 25u) >> 6) & 0x00ff00ffu),            /* 135 - .. - .. */
                                                ((cr2 * 13u) >> 4) & 0x00ff00ffu);
                    if (OUT == KF_OUT_RGBS) {
                        /* sums over my 2 x 4 pixels as int16 pairs {pixels 0, 1 | pixels 2, 3} (both rows): the two 2 x 2 cells
                         * of s = 2; larger cells are sums over neighbouring lanes (q = lane >> 2 down, h = lane & 3 across) */
                        unsigned aR = 0, aG = 0, aB = 0;
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            const unsigned w = r ? y1 : y0;
                            const unsigned te = ((mvg_pair_even(w) * 149u) >> 7) & 0x01ff01ffu;
                            const unsigned to = ((mvg_pair_odd(w) * 149u) >> 7) & 0x01ff01ffu;
                            aR += mvg_add_clip8x2(te, rC) + mvg_add_clip8x2(to, rC);
                            aG += mvg_add_clip8x2(te, gC) + mvg_add_clip8x2(to, gC);
                            aB += mvg_add_clip8x2(te, bC) + mvg_add_clip8x2(to, bC);
                        }
                        const int q = lane >> 2, h = lane & 3;
                        if constexpr (sl == 1) {
                            uint16_t *d = reinterpret_cast<uint16_t *>(s.u.rgb + q * KF_RGBS_PITCH + j * 24 + h * 6);
                            const unsigned R0 = ((aR & 0xffffu) + 2) >> 2, R1 = ((aR >> 16) + 2) >> 2;
                            const unsigned G0 = ((aG & 0xffffu) + 2) >> 2, G1 = ((aG >> 16) + 2) >> 2;
                            const unsigned B0 = ((aB & 0xffffu) + 2) >> 2, B1 = ((aB >> 16) + 2) >> 2;
                            d[0] = (uint16_t)(R0 | (G0 << 8)); d[1] = (uint16_t)(B0 | (R1 << 8)); d[2] = (uint16_t)(G1 | (B1 << 8));
                        } else {
                            unsigned R = (aR & 0xffffu) + (aR >> 16), G = (aG & 0xffffu) + (aG >> 16), B = (aB & 0xffffu) + (aB >> 16);
                            R += __shfl_xor_sync(MVG_FULL, R, 4); G += __shfl_xor_sync(MVG_FULL, G, 4); B += __shfl_xor_sync(MVG_FULL, B, 4);
                            if constexpr (sl >= 3) {
                                R += __shfl_xor_sync(MVG_FULL, R, 8); G += __shfl_xor_sync(MVG_FULL, G, 8); B += __shfl_xor_sync(MVG_FULL, B, 8);
                                R += __shfl_xor_sync(MVG_FULL, R, 1); G += __shfl_xor_sync(MVG_FULL, G, 1); B += __shfl_xor_sync(MVG_FULL, B, 1);
                            }
                            if constexpr (sl == 4) {
                                R += __shfl_xor_sync(MVG_FULL, R, 16); G += __shfl_xor_sync(MVG_FULL, G, 16); B += __shfl_xor_sync(MVG_FULL, B, 16);
                                R += __shfl_xor_sync(MVG_FULL, R, 2); G += __shfl_xor_sync(MVG_FULL, G, 2); B += __shfl_xor_sync(MVG_FULL, B, 2);
                            }
                            /* the first lane of a cell writes its pixel */
                            const bool writer = sl == 2 ? !(q & 1) : sl == 3 ? !(q & 3) && !(h & 1) : lane == 0;
                            if (writer) {
                                constexpr unsigned rnd = 1u << (2 * sl - 1);
                                constexpr int qs = sl > 1 ? sl - 1 : 0, hs = sl > 2 ? sl - 2 : 0;
                                uint8_t *d = s.u.rgb + (q >> qs) * KF_RGBS_PITCH + (j * per + (h >> hs)) * 3;
                                d[0] = (uint8_t)((R + rnd) >> (2 * sl)); d[1] = (uint8_t)((G + rnd) >> (2 * sl)); d[2] = (uint8_t)((B + rnd) >> (2 * sl));
                            }
                        }
                    } else
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        const unsigned w = r ? y1 : y0;
                        const unsigned te = ((mvg_pair_even(w) * 149u) >> 7) & 0x01ff01ffu;       /* pixels 0, 2 */
                        const unsigned to = ((mvg_pair_odd(w) * 149u) >> 7) & 0x01ff01ffu;        /* pixels 1, 3 */
                        const unsigned Re = mvg_add_clip8x2(te, rC), Ro = mvg_add_clip8x2(to, rC);
                        const unsigned Ge = mvg_add_clip8x2(te, gC), Go = mvg_add_clip8x2(to, gC);
                        const unsigned Be = mvg_add_clip8x2(te, bC), Bo = mvg_add_clip8x2(to, bC);
                        const unsigned X = __byte_perm(Re, Ge, 0x6240);       /* R0 G0 R2 G2 */
                        const unsigned Y = __byte_perm(Be, Ro, 0x6240);       /* B0 R1 B2 R3 */
                        const unsigned Z = __byte_perm(Go, Bo, 0x6240);       /* G1 B1 G3 B3 */
                        unsigned *d = KF_WO_GROUP ? reinterpret_cast<unsigned *>(rgb_dst + r * 8 * KF_RGB_GSTRIDE + 48 * j)
                                                  : reinterpret_cast<unsigned *>(rgb_dst + r * KF_RGB_STRIDE + 48 * (j & 1));
                        d[0] = __byte_perm(X, Y, 0x5410);                     /* R0 G0 B0 R1 */
                        d[1] = __byte_perm(Z, X, 0x7610);                     /* G1 B1 R2 G2 */
                        d[2] = __byte_perm(Y, Z, 0x7632);                     /* B2 R3 G3 B3 */
                    }
                }
                /* next macroblock: row -1, x = 15 / 7 becomes x = -1 */
                *cn_dst = *cn_src;
                __syncwarp();
                /* this macroblock's residual is spent: its slot takes macroblock j of the next group */
                if (j < n_next && lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mvg_bulk_load(tile + j * KF_MBS, lv_row + (size_t)(mx + KF_GROUP) * 384, 768u, &s.mbar);
                }
                if (OUT == KF_OUT_RGBS && j == nmb - 1) {
                    /* the group's `per` rows of nmb * per * 3 bytes: 32-bit pieces where rows are whole words (s = 2, 4), bytes
                     * otherwise; piece idx of a full group's geometry, of which a short last group uses the first columns */
                    constexpr int unit = sl <= 2 ? 4 : 1, ppr = KF_GROUP * per * 3 / unit, total = per * ppr;
#pragma unroll 1
                    for (int idx = lane; idx < total; idx += 32) {
                        const int r = idx / ppr, c = idx - r * ppr;
                        if (c * unit < nmb * per * 3) {
                            if (unit == 4) *reinterpret_cast<unsigned *>(wo_run + (size_t)r * opitch + 4 * c) = *reinterpret_cast<const unsigned *>(s.u.rgb + r * KF_RGBS_PITCH + 4 * c);
                            else wo_run[(size_t)r * opitch + c] = s.u.rgb[r * KF_RGBS_PITCH + c];
                        }
                    }
                    wo_run += KF_GROUP * per * 3;
                    __syncwarp();
                }
#if KF_WO_GROUP
                if (OUT == KF_OUT_RGB && j == nmb - 1) {
                    /* the group's 16 rows x 192 bytes (48 per macroblock present) */
                    const int wc = lane & 3, ws = (lane >> 3) + 4 * ((lane >> 2) & 1);     /* piece within 64 bytes, staging row (+ 8 hb) */
                    const uint8_t *rd = s.u.rgb + ws * KF_RGB_GSTRIDE + wc * 16;
                    uint8_t *gd = wo_run + (size_t)(2 * ws) * pitch + wc * 16;
#pragma unroll
                    for (int hb = 0; hb < 2; hb++)
#pragma unroll
                        for (int m = 0; m < 3; m++)
                            if (4 * m + wc < 3 * nmb)
                                *reinterpret_cast<uint4 *>(gd + (size_t)hb * pitch + 64 * m) =
                                    *reinterpret_cast<const uint4 *>(rd + hb * 8 * KF_RGB_GSTRIDE + 64 * m);
                    wo_run += 192;
                    __syncwarp();
                }
#else
                if (OUT == KF_OUT_RGB && ((j & 1) || j == nmb - 1)) {
                    /* the pair's 16 rows x 96 bytes (48 for a lone last macroblock): 16-byte chunks in row-major order over
                     * the lanes, so that a pair leaves as whole 32-byte sectors (96 bytes per row at a multiple of 96) */
                    const int n_here = (j & 1) + 1;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const int ch = lane + 32 * k, r = ch / 6, col = ch - r * 6;
                        if (col < 3 * n_here) {
                            const uint4 v = *reinterpret_cast<const uint4 *>(s.u.rgb + r * KF_RGB_STRIDE + col * 16);
#ifdef KF_RGB_LINEAR    /* DEV, timing only (wrong picture): the pair's 1536 bytes in one piece instead of 16 row segments */
                            *reinterpret_cast<uint4 *>(p.rgb + (size_t)slot * ((size_t)n_mb * 768) + ((size_t)row * W + mx - (j & 1)) * 768 + ch * 16) = v;
#else
                            *reinterpret_cast<uint4 *>(wo_run + (size_t)r * pitch + col * 16) = v;
#endif
                        }
                    }
                    wo_run += 96;
                    __syncwarp();
                }
#endif
            }
        }
        KF_STAT(st[6] += clock64() - trow;)
    }
    KF_STAT(if (lane == 0 && p.stats) for (int i = 0; i < 8; i++) atomicAdd(p.stats + i, st[i]);)
#ifdef MVG_CHECKED
    __syncwarp();
    if (lane < 4) MVG_ASSERT(s.c0[lane] == MVG_CANARY_WORD && s.c1[lane] == MVG_CANARY_WORD && s.c2[lane] == MVG_CANARY_WORD && s.c3[lane] == MVG_CANARY_WORD, 2);
#endif
}
