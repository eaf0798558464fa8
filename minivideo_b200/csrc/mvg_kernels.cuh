/*
 * mvg_kernels.cuh -- the three sm_100a kernels of the intra reconstruction path.
 *
 *   k1_dequant_idct   dequantisation + 4x4/8x8 inverse integer transforms,
 *                     Intra16x16 luma-DC Hadamard, chroma-DC 2x2        (HBM-bound)
 *   k2_wavefront      Intra4x4/8x8/16x16 + chroma prediction, residual add,
 *                     macroblock-row wavefront batched over pictures    (dependency-bound)
 *   k3_rgb            fused 4:2:0 -> RGB24 convert (+ box downscale)    (HBM-bound)
 *
 * Arithmetic follows the reference bit for bit (citations: minivideo/src/decoder/h264/
 * in the reference tree, which is not part of this repository):
 *   h264_transform.c        -> k1 (quant4x4 :1100, idct4x4 :1145, quant8x8 :1256,
 *                              idct8x8 :1295, lumadc :756, chromadc :827-936)
 *   h264_intra_prediction.c -> k2 (4x4 :315-926, 8x8 :1107-1793, 16x16 :1809-2141,
 *                              chroma :2157-2564) + residual add h264_transform.c:152,219,267,393
 *   export_utils.c:209-324  -> k3
 * No tensor cores: none of this is a dense contraction.
 */
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "mvg_internal.h"

#define MVG_FULL 0xffffffffu

/* ---- checked build (-DMVG_CHECKED, tests/tools/checked_build.py): the stand-in for compute-sanitizer, which is closed
 * on this pool.  Every computed shared-memory address of the prediction stage and every list index of the transform
 * stage is compared with the bounds of the warp's own record, the records carry canary words between their members,
 * and violations are counted in a device variable that mvg_debug_check_kernels() reads.  Nothing of this exists in
 * the product build. */
#ifdef MVG_CHECKED
__device__ unsigned long long mvg_check_fail[8];    /* [0] address outside the warp record, [1] list index, [2] canary, [3] table index */
#define MVG_ASSERT(cond, slot) do { if (!(cond)) atomicAdd(&mvg_check_fail[slot], 1ull); } while (0)
#define MVG_CANARY(name) uint32_t name[4];
#define MVG_CANARY_WORD 0xC0DEC0DEu
#else
#define MVG_ASSERT(cond, slot) do { } while (0)
#define MVG_CANARY(name)
#endif

/* ------------------------------------------------------------------------- */
/* constant tables                                                             */

struct MvgTables {
    int32_t ls4[3][6][16];      /* LevelScale4x4[c][q][i*4+j] */
    int32_t ls4q[3][52][16];    /* per qP: LevelScale4x4[c][qP%6] << (qP/6-4) when qP >= 24, else unshifted */
    int32_t ls8[6][64];         /* LevelScale8x8[0][q][i*8+j] */
    uint8_t zz8inv[64];         /* zz8inv[row*8+col] = zig-zag index k of that position */
    int32_t cb_qp_offset, cr_qp_offset;
};

__device__ __forceinline__ int mvg_clip8(int v) { return min(max(v, 0), 255); }

/* lane index read once through an opaque instruction: from `threadIdx.x & 31` the compiler re-reads the thread id
 * (S2R, tens of cycles) inside the loops whenever it is short of registers */
__device__ __forceinline__ int mvg_lane()
{
    int l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

/* A per-lane INTEGER constant computed once in a kernel's prologue, made opaque so that the compiler keeps it in a
 * register instead of re-deriving it from the lane index inside the row loop.  Only for integers (offsets): a pointer
 * that goes through an asm loses its address space, and every access through it becomes a generic load or store. */
__device__ __forceinline__ int mvg_keep(int v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ unsigned mvg_keep(unsigned v) { asm volatile("" : "+r"(v)); return v; }

/* Table 8-15: QPC as a function of qPI >= 30 (h264_transform.c:71) */
__constant__ unsigned char mvg_qpc_tab[22] = {29,30,31,32,32,33,34,34,35,35,36,36,37,37,37,38,38,38,39,39,39,39};

/* h264_transform.c:598-637 (8-bit video: QpBdOffsetC = 0) */
__device__ __forceinline__ int mvg_chroma_qp(int qp_y, int offset)
{
    const int qpi = min(max(qp_y + offset, 0), 51);
    return qpi < 30 ? qpi : (int)mvg_qpc_tab[qpi - 30];
}

/* ========================================================================= */
/* Kernel 0: packed levels -> dense levels (only on the end-to-end path; see mvgpu.h) */

struct K0Params {
    const uint32_t *nz_blocks, *word_off;   /* [n_mbs]                                    */
    const uint64_t *pic_base;               /* [n_pics + 1] first word of each picture in `words`, then the end */
    const uint16_t *words;
    int16_t        *coeff;                  /* [n_mbs][384]                               */
    long long       n_mbs;
    int             mbs_per_pic;
};

/* One warp per macroblock, lane b < 24 = chunk b: its mask sits at rank(b) among the coded chunks, its
 * levels start after all masks plus the levels of the chunks before it (warp prefix sum of popcounts). */
__global__ void __launch_bounds__(256)
k0_expand_levels(K0Params p)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long mb = warp0; mb < p.n_mbs; mb += n_warps) {
        const unsigned nzb = __ldg(p.nz_blocks + mb) & 0x00FFFFFFu;
        const long long pic = mb / p.mbs_per_pic;
        const uint16_t *w = p.words + __ldg(p.pic_base + pic) + __ldg(p.word_off + mb);
        const uint16_t *end = p.words + __ldg(p.pic_base + pic + 1);        /* a malformed batch must not read past its picture */
        const bool coded = (nzb >> lane) & 1u;
        const uint16_t *mp = w + __popc(nzb & ((1u << lane) - 1u));
        const unsigned mask = (coded && mp < end) ? (unsigned)__ldg(mp) : 0u;
        int pre = __popc(mask);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(MVG_FULL, pre, o);
            if (lane >= o) pre += t;
        }
        const uint16_t *lv = w + __popc(nzb) + pre - __popc(mask);
        unsigned out[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        unsigned m = mask;
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const unsigned v = lv < end ? (unsigned)__ldg(lv) : 0u;
            lv++;
#pragma unroll
            for (int q = 0; q < 8; q++)
                if ((k >> 1) == q) out[q] |= v << (16 * (k & 1));
        }
        if (lane < 24) {
            uint4 *dst = reinterpret_cast<uint4 *>(p.coeff + mb * 384 + lane * 16);
            dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
            dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
        }
    }
}

/* ========================================================================= */
/* Kernel 1                                                                    */

struct K1Params {
    const uint8_t *mb_kind, *i16_mode, *chroma_mode, *luma_modes;
    const int8_t  *qp_y;
    const int16_t *coeff;       /* [n_mbs][384] */
    int16_t       *resid;       /* [n_mbs][384] */
    MvgMbCtl      *ctl;         /* [n_mbs]      */
    const MvgTables *tab;
    long long      n_mbs;
};

/* spec 8.5.12.2 / h264_transform.c:1145-1191, one 4-point butterfly */
__device__ __forceinline__ void mvg_bfly4(int a, int b, int c, int d, int &o0, int &o1, int &o2, int &o3)
{
    int e0 = a + c, e1 = a - c, e2 = (b >> 1) - d, e3 = b + (d >> 1);
    o0 = e0 + e3; o1 = e1 + e2; o2 = e1 - e2; o3 = e0 - e3;
}

/* spec 8.5.13.2 / h264_transform.c:1308-1378, one 8-point pass */
__device__ __forceinline__ void mvg_idct8_1d(int (&v)[8])
{
    int a0 = v[0] + v[4];
    int a1 = -v[3] + v[5] - v[7] - (v[7] >> 1);
    int a2 = v[0] - v[4];
    int a3 = v[1] + v[7] - v[3] - (v[3] >> 1);
    int a4 = (v[2] >> 1) - v[6];
    int a5 = -v[1] + v[7] + v[5] + (v[5] >> 1);
    int a6 = v[2] + (v[6] >> 1);
    int a7 = v[3] + v[5] + v[1] + (v[1] >> 1);
    int b0 = a0 + a6, b1 = a1 + (a7 >> 2), b2 = a2 + a4, b3 = a3 + (a5 >> 2);
    int b4 = a2 - a4, b5 = (a3 >> 2) - a5, b6 = a0 - a6, b7 = a7 - (a1 >> 2);
    v[0] = b0 + b7; v[1] = b2 + b5; v[2] = b4 + b3; v[3] = b6 + b1;
    v[4] = b6 - b1; v[5] = b4 - b3; v[6] = b2 - b5; v[7] = b0 - b7;
}

/* two int32 -> saturated int16 pair (hi:lo) in one instruction */
__device__ __forceinline__ unsigned mvg_pack_sat16(int hi, int lo)
{
    unsigned d;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(hi), "r"(lo));
    return d;
}
/* {clamp(hi >> 6, -512, 511), clamp(lo >> 6, -512, 511)} as an int16 pair: saturate to int16 first
 * (>> is monotonic, so clamp and shift commute), then shift both halves at once and sign-extend the two
 * 10-bit fields with a packed subtract.  The residual is clamped because kernel 2 adds it to the
 * prediction with packed 16-bit arithmetic: Clip1(pred + r) does not change for any clamp range that
 * contains [-255, 255], and 255 + 511 cannot wrap. */
__device__ __forceinline__ unsigned mvg_pack_shr6(int hi, int lo)
{
    const unsigned t = ((mvg_pack_sat16(hi, lo) >> 6) & 0x03ff03ffu) ^ 0x02000200u;
    return __vsub2(t, 0x02000200u);
}

#define K1_WARPS 12         /* warps per CTA; two CTAs per SM     */
#define K1_GROUP 4          /* macroblocks per warp iteration     */
#define K1_TILE  (K1_GROUP * 384)

/* ---- bulk asynchronous copies (TMA, 1-D) and their mbarrier ------------------- */
__device__ __forceinline__ uint32_t mvg_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mvg_mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mvg_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mvg_mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mvg_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mvg_mbar_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(mvg_smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mvg_bulk_load(void *dst_smem, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(mvg_smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(mvg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mvg_bulk_store(void *dst, const void *src_smem, unsigned bytes)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(dst), "r"(mvg_smem_u32(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void mvg_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

/* per-warp scratch of the transform stage for a group of G macroblocks */
template <int G>
struct MvgXfScratch {
    int32_t  dc[G][24];                         /* dequantised DC of Intra16x16 luma / chroma blocks   */
    int32_t  f1[G][16];                         /* first stage of the luma DC Hadamard                 */
    int32_t  tr[4][8][9];                       /* 8x8 transpose, padded                               */
    uint32_t meta[G];                           /* mb_kind | QPY << 8 | (shift | rounding << 8) of quant4x4 at QPY << 16 */
    int32_t  ls0[G];                            /* LevelScale4x4 (luma, position 0) at QPY, shift folded in: DC-only luma blocks */
    uint8_t  list4[G * 24 + 8];                 /* 4x4 blocks that need the full transform             */
    uint8_t  list8[G * 4 + 4];                  /* 8x8 blocks with non-zero levels                     */
};

/* dequantisation tables in shared memory (filled by mvg_xf_load_tables) */
struct MvgXfTables {
    int32_t  ls4[3 * 6 * 16];
    int32_t  ls4q[3 * 52 * 16];                 /* per qP; << (qP/6-4) folded in when qP >= 24 */
    int32_t  ls8[6 * 64];
    __align__(8) uint8_t zz8inv[64];
    uint8_t  dcsh[52];                          /* 4 - qP / 6 below qP 24, else 0 */
    uint16_t qpc[2][52];                        /* QPC | QPC / 6 << 8 for Cb, Cr by QPY (derivChromaQP) */
};

struct K1WarpSmem {
    __align__(128) int16_t tile[2][K1_TILE];    /* levels in, residual out (in place), double buffered */
    MvgXfScratch<K1_GROUP> x;
    __align__(8) uint64_t mbar[2];
};

/* position (bx,by) -> luma4x4BlkIdx (h264_spatial.c:210-225 inverted) */
__device__ __forceinline__ int mvg_blk_of(int bx, int by) { return (bx & 1) | ((by & 1) << 1) | ((bx >> 1) << 2) | ((by >> 1) << 3); }

/* cooperative fill of the dequantisation tables (all threads of the CTA; the caller synchronises) */
__device__ __forceinline__ void mvg_xf_load_tables(MvgXfTables &t, const MvgTables *tab)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < 3 * 6 * 16; i += nt) t.ls4[i] = (&tab->ls4[0][0][0])[i];
    for (int i = tid; i < 3 * 52 * 16; i += nt) t.ls4q[i] = (&tab->ls4q[0][0][0])[i];
    for (int i = tid; i < 6 * 64; i += nt) t.ls8[i] = (&tab->ls8[0][0])[i];
    for (int i = tid; i < 64; i += nt) t.zz8inv[i] = tab->zz8inv[i];
    for (int i = tid; i < 52; i += nt) t.dcsh[i] = (uint8_t)(i > 23 ? 0 : 4 - i / 6);
    for (int i = tid; i < 104; i += nt) {
        const int pl = i >= 52, qpc = mvg_chroma_qp(i - 52 * pl, pl ? tab->cr_qp_offset : tab->cb_qp_offset);
        t.qpc[pl][i - 52 * pl] = (uint16_t)(qpc | ((qpc / 6) << 8));
    }
}

/* Side information of a group of G <= 4 macroblocks, one 32-bit word per lane: lane = 8*j + t for macroblock j;
 * t = 0..3 luma modes 4t..4t+3, t = 4 mb_kind, 5 QPY, 6 Intra16x16PredMode, 7 intra_chroma_pred_mode.
 * `first` = index of the group's first macroblock, `n` = macroblocks in the group. */
struct MvgSideInfo {
    const uint8_t *src;         /* this lane's array, at macroblock 0 */
    int stride;                 /* bytes per macroblock in that array */
    __device__ __forceinline__ void init(int lane, const uint8_t *mb_kind, const uint8_t *i16_mode, const uint8_t *chroma_mode,
                                         const uint8_t *luma_modes, const int8_t *qp_y)
    {
        const int mt = lane & 7;
        src = mt < 4 ? luma_modes + 4 * mt : mt == 4 ? mb_kind : mt == 5 ? reinterpret_cast<const uint8_t *>(qp_y)
                     : mt == 6 ? i16_mode : chroma_mode;
        stride = mt < 4 ? 16 : 1;
    }
    __device__ __forceinline__ unsigned load(int lane, long long first, int n) const
    {
        unsigned v = 0u;
        if ((lane >> 3) < n) {
            const uint8_t *q = src + (first + (lane >> 3)) * stride;
            if ((lane & 7) < 4) v = __ldg(reinterpret_cast<const unsigned *>(q));
            else v = (unsigned)__ldg(q);
        }
        return v;
    }
};

/* The transform stage for a group of G macroblocks whose levels sit in `tile` (G x 384 int16, zig-zag order as in
 * mvgpu.h); the residual replaces them in place, block-major (24 blocks x 16 int16 per macroblock, see MvgMbCtl):
 *   - Intra16x16 luma DC (4x4 Hadamard, h264_transform.c:756-812) and chroma DC (2x2, :827-936)
 *     are done first, separably, a few lanes per macroblock;
 *   - every 4x4 block is classified: all-zero (nothing to do), DC-only (every residual sample is
 *     (d00 + 32) >> 6, spec 8.5.12.2 with a single non-zero input) or general; the general blocks
 *     of the whole group are compacted with ballots and run through quant4x4 + idct4x4
 *     (h264_transform.c:1100-1191) 32 at a time, one block per lane, entirely in registers;
 *   - non-zero 8x8 blocks (Intra8x8 luma, quant8x8/idct8x8 :1256-1383) are compacted the same
 *     way and transformed 4 per pass, 8 lanes per block, transposed through shared memory.
 * `meta` = the group's side information (MvgSideInfo), nmb = macroblocks present (1..G).  Warp-collective. */
template <int G, int MBS = 384>
__device__ __forceinline__ void mvg_xf_group(int16_t *tile, MvgXfScratch<G> &s, const MvgXfTables &T, unsigned meta, int nmb, int lane)
{
    /* MBS = int16 per macroblock slot of `tile`: 384 when the group is one contiguous piece (kernel 1: it leaves with one
     * bulk store), 392 in the fused kernel -- 16 bytes of padding per macroblock put the DC levels of the four macroblocks
     * (every 32 bytes: banks 0, 8, 16, 24 only) into different banks */
    const int mj = lane >> 3, mt = lane & 7;
    /* per-lane view of "my" macroblock j = lane >> 3 */
    const int kind_j = (int)__shfl_sync(MVG_FULL, meta, (lane & ~7) + 4);
    /* QPY outside 0..51 cannot come out of a conforming parse; clamp so that a bad batch cannot index past the tables */
    const int qp_j = min(max((int)(signed char)__shfl_sync(MVG_FULL, meta, (lane & ~7) + 5), 0), 51);
    if (mt == 0 && mj < G) {
        /* what the classification pass needs of a macroblock besides its kind: one word and the DC scale, instead of
         * three dependent table reads per block */
        const unsigned sh = T.dcsh[qp_j];
        s.meta[mj] = (unsigned)kind_j | ((unsigned)(qp_j & 255) << 8) | (sh << 16) | (((1u << sh) >> 1) << 24);
        s.ls0[mj] = T.ls4q[qp_j * 16];
    }

    /* ---------------- DC transforms ---------------- */
    if (mj < nmb) {
        const int16_t *cf = tile + mj * MBS;
        if (kind_j == MVG_MB_I16x16 && mt < 4) {         /* row mt of c: t = c * H */
            const int a = cf[mvg_blk_of(0, mt) * 16], b = cf[mvg_blk_of(1, mt) * 16];
            const int c = cf[mvg_blk_of(2, mt) * 16], d = cf[mvg_blk_of(3, mt) * 16];
            int32_t *f1 = s.f1[mj] + mt * 4;
            f1[0] = a + b + c + d; f1[1] = a + b - c - d; f1[2] = a - b - c + d; f1[3] = a - b + c - d;
        } else if (mt == 4 || mt == 5) {                 /* chroma plane mt-4: f = A c A, then scale */
            const int pl = mt - 4;
            const int16_t *cc = cf + 256 + pl * 64;
            const int c00 = cc[0], c01 = cc[16], c10 = cc[32], c11 = cc[48];
            const int qe = T.qpc[pl][qp_j], qpc = qe & 255, qd = qe >> 8;
            const int ls00 = T.ls4[((pl + 1) * 6 + (qpc - 6 * qd)) * 16];
            const int f[4] = {c00 + c01 + c10 + c11, c00 - c01 + c10 - c11, c00 + c01 - c10 - c11, c00 - c01 - c10 + c11};
#pragma unroll
            for (int k = 0; k < 4; k++) s.dc[mj][16 + pl * 4 + k] = (int)((unsigned)(f[k] * ls00) << qd) >> 5;
        }
    }
    __syncwarp();
    if (mj < nmb && kind_j == MVG_MB_I16x16 && mt < 4) { /* column mt: f = H * t, then scale */
        const int32_t *f1 = s.f1[mj];
        const int a = f1[mt], b = f1[4 + mt], c = f1[8 + mt], d = f1[12 + mt];
        const int f[4] = {a + b + c + d, a + b - c - d, a - b - c + d, a - b + c - d};
        const int qd = qp_j / 6, ls00 = T.ls4[(qp_j - 6 * qd) * 16];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int t = f[i] * ls00;
            s.dc[mj][mvg_blk_of(mt, i)] = (qp_j >= 36) ? (int)((unsigned)t << (qd - 6)) : ((t + (1 << (5 - qd))) >> (6 - qd));
        }
    }
    __syncwarp();

    /* ---------------- classify the 4x4 blocks, compact the general ones ---------------- */
    int n4 = 0, n8 = 0;
#pragma unroll 1        /* code size: the fused kernel has to fit the instruction cache */
    for (int r = 0; r < (G * 24 + 31) / 32; r++) {
        /* u / 24 for u = lane + 32 r < 96: r, and one more from lane 24 - 8 r on */
        const int u = lane + 32 * r, j0 = G <= 4 ? r + (lane >= 24 - 8 * r) : u / 24, b = u - 24 * j0;
        /* straight-line code: every lane loads a block (its own, or block b of macroblock 0 beyond the last
         * macroblock of a short group) and derives all three answers; only the DC-only rewrite is conditional */
        const bool live = j0 < nmb;
        const int j = live ? j0 : 0;
        const unsigned mw = s.meta[j];
        const int kind = mw & 255;
        const bool is8 = kind == MVG_MB_I8x8 && b < 16;
        uint4 *blk = reinterpret_cast<uint4 *>(tile + j * MBS + b * 16);
        /* a lane's block is two 16-byte halves 32 bytes apart from its neighbour's: read in the same order by all lanes,
         * the eight lanes of a quarter warp touch only four of the eight 16-byte bank groups (two wavefronts per quarter).
         * Lanes 4..7 of every eight take the second half first: 2 l + (l >> 2 & 1) covers all eight groups. */
        const int o = (lane >> 2) & 1;
        const uint4 wa = blk[o], wb = blk[o ^ 1];
        const unsigned x0 = o ? wb.x : wa.x, x4 = o ? wa.x : wb.x;        /* first word of the first / second half */
        const unsigned rest = (x0 & 0xffff0000u) | x4 | wa.y | wa.z | wa.w | wb.y | wb.z | wb.w;
        const int dcraw = (short)(x0 & 0xffff);
        const bool nz8q = live && is8 && (rest | (x0 & 0xffffu)) != 0;
        const bool general = live && !is8 && rest != 0;
        {
            /* DC only: every residual sample is (d00 + 32) >> 6.  d00 = c00 (already dequantised by the DC
             * transforms, h264_transform.c:1126-1129) for chroma and Intra16x16, else quant4x4 of the level:
             * (c * LS + rnd) >> sh with sh = 0 from qP 24 on (the left shift is folded into ls4q) */
            const bool has_dc = b >= 16 || kind == MVG_MB_I16x16;
            const int sh = (mw >> 16) & 255;
            const int plain = (dcraw * s.ls0[j] + (int)(mw >> 24)) >> sh;
            const int d = has_dc ? s.dc[j][b] : plain;
            const int rv = min(max((d + 32) >> 6, -512), 511);
            if (live && !is8 && rest == 0 && (rv != 0 || dcraw != 0)) {
                const unsigned pk = (unsigned)(rv & 0xffff) * 0x10001u;
                blk[o] = make_uint4(pk, pk, pk, pk); blk[o ^ 1] = make_uint4(pk, pk, pk, pk);
            }
        }
        const unsigned gb = __ballot_sync(MVG_FULL, general);
        if (general) s.list4[n4 + __popc(gb & ((1u << lane) - 1))] = (uint8_t)u;
        n4 += __popc(gb);
        /* Intra8x8: slots 4*b8..4*b8+3 are the four quarters of 8x8 block b8 (aligned lane quads) */
        const unsigned qb = __ballot_sync(MVG_FULL, nz8q);
        const bool any8 = live && is8 && ((qb >> (lane & ~3)) & 0xFu) != 0;
        const bool lead8 = any8 && (lane & 3) == 0;
        const unsigned lb = __ballot_sync(MVG_FULL, lead8);
        if (lead8) s.list8[n8 + __popc(lb & ((1u << lane) - 1))] = (uint8_t)(j * 4 + (b >> 2));
        n8 += __popc(lb);
    }
    __syncwarp();

    /* ---------------- general 4x4 blocks, 32 per pass ---------------- */
    for (int base = 0; base < n4; base += 32) {
        if (base + lane < n4) {
            const int u = s.list4[base + lane], j = u / 24, b = u - 24 * j;
            MVG_ASSERT(base + lane < G * 24 + 8 && j < nmb, 1);
            const unsigned mw = s.meta[j];
            const int kind = mw & 255, qp = (signed char)(mw >> 8);
            const int comp = b < 16 ? 0 : (b < 20 ? 1 : 2);
            const int qpb = comp ? (T.qpc[comp - 1][qp] & 255) : qp;
            uint4 *blk = reinterpret_cast<uint4 *>(tile + j * MBS + b * 16);
            const uint4 a = blk[0], bb = blk[1];
            int c[16];                      /* zig-zag k -> (row,col): utils.h:64 / spec Table 8-13 */
            c[0] = (short)(a.x & 0xffff); c[1] = (int)a.x >> 16; c[4] = (short)(a.y & 0xffff); c[8] = (int)a.y >> 16;
            c[5] = (short)(a.z & 0xffff); c[2] = (int)a.z >> 16; c[3] = (short)(a.w & 0xffff); c[6] = (int)a.w >> 16;
            c[9] = (short)(bb.x & 0xffff); c[12] = (int)bb.x >> 16; c[13] = (short)(bb.y & 0xffff); c[10] = (int)bb.y >> 16;
            c[7] = (short)(bb.z & 0xffff); c[11] = (int)bb.z >> 16; c[14] = (short)(bb.w & 0xffff); c[15] = (int)bb.w >> 16;
            const bool keep_dc = comp != 0 || kind == MVG_MB_I16x16;
            const int dc_in = keep_dc ? s.dc[j][b] : 0;
            /* quant4x4, h264_transform.c:1100-1134 */
            const int4 *lq = reinterpret_cast<const int4 *>(T.ls4q + (comp * 52 + qpb) * 16);
            const int4 l0 = lq[0], l1 = lq[1], l2 = lq[2], l3 = lq[3];
            const int ls[16] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w, l2.x, l2.y, l2.z, l2.w, l3.x, l3.y, l3.z, l3.w};
            {   /* one code path for both halves of h264_transform.c:1112-1123: from qP 24 on the left shift is part of
                 * ls4q and what remains is (c * LS + 0) >> 0 */
                const int sh = T.dcsh[qpb], rnd = (1 << sh) >> 1;
#pragma unroll
                for (int k = 0; k < 16; k++) c[k] = (c[k] * ls[k] + rnd) >> sh;
            }
            if (keep_dc) c[0] = dc_in;
            c[0] += 32;                     /* rounding of the final >> 6 (h264_transform.c:1190) */
#pragma unroll
            for (int i = 0; i < 4; i++)     /* idct4x4: rows, then columns */
                mvg_bfly4(c[4 * i], c[4 * i + 1], c[4 * i + 2], c[4 * i + 3], c[4 * i], c[4 * i + 1], c[4 * i + 2], c[4 * i + 3]);
#pragma unroll
            for (int q = 0; q < 4; q++)
                mvg_bfly4(c[q], c[4 + q], c[8 + q], c[12 + q], c[q], c[4 + q], c[8 + q], c[12 + q]);
            uint4 o0, o1;
            o0.x = mvg_pack_shr6(c[1], c[0]);   o0.y = mvg_pack_shr6(c[3], c[2]);
            o0.z = mvg_pack_shr6(c[5], c[4]);   o0.w = mvg_pack_shr6(c[7], c[6]);
            o1.x = mvg_pack_shr6(c[9], c[8]);   o1.y = mvg_pack_shr6(c[11], c[10]);
            o1.z = mvg_pack_shr6(c[13], c[12]); o1.w = mvg_pack_shr6(c[15], c[14]);
            blk[0] = o0; blk[1] = o1;
        }
    }

    /* ---------------- non-zero 8x8 blocks, 4 per pass ---------------- */
    for (int base = 0; base < n8; base += 4) {
        const int slot = base + (lane >> 3), row = lane & 7;
        const bool act = slot < n8;
        int v[8];
        int16_t *o8 = tile;
        if (act) {
            const int id = s.list8[slot], j = id >> 2, b8 = id & 3;
            MVG_ASSERT(slot < G * 4 + 4 && j < nmb, 1);
            const int qp = (signed char)(s.meta[j] >> 8);
            const int qd8 = qp / 6;
            const int32_t *l8 = T.ls8 + (qp - 6 * qd8) * 64 + row * 8;
            const int16_t *in = tile + j * MBS + b8 * 64;
            const uint2 zz = *reinterpret_cast<const uint2 *>(T.zz8inv + row * 8);   /* scan positions of my row */
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const unsigned k = ((q < 4 ? zz.x : zz.y) >> (8 * (q & 3))) & 255u;
                v[q] = (int)in[k] * l8[q];                                  /* quant8x8, h264_transform.c:1256-1284 */
            }
            {   /* both halves of h264_transform.c:1268-1279 as (v * 2^up + rnd) >> down: up = qP/6 - 6 and rnd = down = 0
                 * from qP 36 on, else up = 0 (the product wraps exactly like the reference's left shift) */
                const int up = max(qd8 - 6, 0), down = max(6 - qd8, 0), rnd = (1 << down) >> 1, mul = 1 << up;
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = (v[q] * mul + rnd) >> down;
            }
            if (row == 0) v[0] += 32;                                       /* rounding of the final >> 6 (:1382) */
            mvg_idct8_1d(v);                                                /* row pass */
#pragma unroll
            for (int q = 0; q < 8; q++) s.tr[lane >> 3][row][q] = v[q];
            o8 = tile + j * MBS + (b8 * 4 + (row >> 2)) * 16 + (row & 3);
        }
        __syncwarp();
        if (act) {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = s.tr[lane >> 3][i][row];     /* this lane now owns column `row` */
            mvg_idct8_1d(v);                                                /* column pass */
#pragma unroll
            for (int i = 0; i < 8; i++) o8[(i >> 2) * 32 + (i & 3) * 4] = (int16_t)min(max(v[i] >> 6, -512), 511);
        }
        __syncwarp();
    }
    __syncwarp();
}

/* the 16-byte control record of a macroblock (MvgMbCtl) from the side-information words of its lanes 8j..8j+7;
 * valid in lane 8j */
__device__ __forceinline__ uint4 mvg_ctl_from_meta(unsigned meta)
{
    /* 4 mode bytes -> 4 nibbles; lane 8j collects its macroblock's record */
    const unsigned nib = (meta & 0xF) | ((meta >> 4) & 0xF0) | ((meta >> 8) & 0xF00) | ((meta >> 12) & 0xF000);
    const unsigned n1 = __shfl_down_sync(MVG_FULL, nib, 1), n2 = __shfl_down_sync(MVG_FULL, nib, 2);
    const unsigned n3 = __shfl_down_sync(MVG_FULL, nib, 3);
    const unsigned k4 = __shfl_down_sync(MVG_FULL, meta, 4), k6 = __shfl_down_sync(MVG_FULL, meta, 6);
    const unsigned k7 = __shfl_down_sync(MVG_FULL, meta, 7);
    return make_uint4((k4 & 255) | ((k6 & 255) << 8) | ((k7 & 255) << 16), nib | (n2 << 16),
                      __funnelshift_l(n1 | (n3 << 16), n1 | (n3 << 16), 8), 0u);
}

/* Kernel 1 (kept as the verification tap behind mvg_download_residual() and as the first stage of the split
 * pipeline): one warp transforms K1_GROUP macroblocks per iteration, in place in shared memory.  The 3 KB of
 * levels arrive with ONE bulk asynchronous copy (cp.async.bulk + mbarrier), the next group's copy is in flight
 * while this one is processed, and the residual leaves with one bulk store -- no per-thread global loads/stores
 * for the payload.  Residual layout out: per macroblock 24 blocks x 16 int16, block-major (see MvgMbCtl). */
__global__ void __launch_bounds__(K1_WARPS * 32, 2)
k1_dequant_idct(K1Params p)
{
    __shared__ MvgXfTables s_tab;
    extern __shared__ __align__(128) uint8_t k1_smem[];         /* K1WarpSmem x K1_WARPS (dynamic: above the 48 KB static limit) */
    K1WarpSmem *s_warp = reinterpret_cast<K1WarpSmem *>(k1_smem);
    mvg_xf_load_tables(s_tab, p.tab);

    const int lane = mvg_lane();
    K1WarpSmem &s = s_warp[threadIdx.x >> 5];
    if (lane == 0) { mvg_mbar_init(&s.mbar[0], 1); mvg_mbar_init(&s.mbar[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const long long n_groups = (p.n_mbs + K1_GROUP - 1) / K1_GROUP;
    const long long stride = (long long)gridDim.x * K1_WARPS;
    long long g = (long long)blockIdx.x * K1_WARPS + (threadIdx.x >> 5);

    MvgSideInfo side;
    side.init(lane, p.mb_kind, p.i16_mode, p.chroma_mode, p.luma_modes, p.qp_y);
    auto issue_load = [&](int buf, long long grp) {
        const long long mb0 = grp * K1_GROUP;
        const unsigned bytes = (unsigned)min((long long)K1_GROUP, p.n_mbs - mb0) * 768u;
        mvg_mbar_expect_tx(&s.mbar[buf], bytes);
        mvg_bulk_load(s.tile[buf], p.coeff + mb0 * 384, bytes, &s.mbar[buf]);
    };

    unsigned nmeta = 0;
    if (g < n_groups) {
        if (lane == 0) issue_load(0, g);
        nmeta = side.load(lane, g * K1_GROUP, (int)min((long long)K1_GROUP, p.n_mbs - g * K1_GROUP));
    }
    unsigned parity = 0;            /* bit b: phase parity of mbar[b] */

    for (int it = 0; g < n_groups; g += stride, it++) {
        const int buf = it & 1;
        const unsigned meta = nmeta;
        {   /* next group: its buffer was the source of the previous bulk store */
            const long long gn = g + stride;
            if (gn < n_groups) {
                if (lane == 0) { mvg_bulk_wait_read(); issue_load(buf ^ 1, gn); }
                nmeta = side.load(lane, gn * K1_GROUP, (int)min((long long)K1_GROUP, p.n_mbs - gn * K1_GROUP));
            }
        }
        const long long mb0 = g * K1_GROUP;
        const int nmb = (int)min((long long)K1_GROUP, p.n_mbs - mb0);
        int16_t *tile = s.tile[buf];

        mvg_mbar_wait(&s.mbar[buf], (parity >> buf) & 1u);
        parity ^= 1u << buf;

        mvg_xf_group<K1_GROUP>(tile, s.x, s_tab, meta, nmb, lane);

        /* ---------------- residual out (one bulk store) + control records ---------------- */
        if (lane == 0) mvg_bulk_store(p.resid + mb0 * 384, tile, (unsigned)nmb * 768u);
        {
            const uint4 rec = mvg_ctl_from_meta(meta);
            if ((lane & 7) == 0 && (lane >> 3) < nmb) *reinterpret_cast<uint4 *>(p.ctl + mb0 + (lane >> 3)) = rec;
        }
        __syncwarp();
    }
    if (lane == 0) mvg_bulk_wait_read();
}

/* ========================================================================= */
/* Kernel 2                                                                    */

struct K2Params {
    const int16_t  *resid;      /* [slot][n_mb][384]                          */
    const MvgMbCtl *ctl;        /* [slot][n_mb]                               */
    uint8_t        *tiles;      /* [slot][n_mb][384] reconstructed macroblocks: 16x16 Y, 8x8 Cb, 8x8 Cr rasters */
    uint2          *halo;       /* [slot][h_mbs][w_mbs][8]: bottom sample row of every MB,
                                   4 data bytes + 4 flag bytes per 64-bit word  */
    int            *work;       /* work counter of this launch (starts at 0)  */
    const MvgLuts  *luts;
    unsigned        epoch;      /* flag value that marks halo words of THIS launch */
    int w_mbs, h_mbs, first_slot, n_pics, group;
    unsigned long long *stats;  /* DEV (-DMVG_K2_PROFILE): cycle accounting, 16 counters */
    unsigned        sel[4];     /* 1 << 8k: dot-product selectors of byte k; kernel parameters so that they are
                                   constant-bank operands instead of per-use uniform moves */
};

#ifdef MVG_K2_PROFILE
#define K2_PROF(...) __VA_ARGS__
#else
#define K2_PROF(...)
#endif
#ifndef K2_POLL_NS
#define K2_POLL_NS   100        /* first sleep of a row that has caught up with the row above; doubles up to 8x */
#endif
#ifndef K2_WARPS
#define K2_WARPS     16         /* warps per CTA; two CTAs per SM                              */
#endif
#define K2_CTL_CHUNK 32         /* control records per chunk (one 16-byte copy per lane)       */
#define K2_RING      4          /* residual ring slots: the pair in use and the pair in flight     */
#define K2_TO(x, y)  (((y) + 1) * MVG_LT_STRIDE + MVG_LT_XOFF + (x))    /* luma tile offset of sample (x, y)   */
#define K2_CO(x, y)  (((y) + 1) * MVG_CT_STRIDE + MVG_CT_XOFF + (x))    /* chroma tile offset of sample (x, y) */

/* Member order matters to the Intra4x4 steps: lanes without a block in a step still run it and read up to 64 bytes
 * below resid[] and 140 bytes below lt[] (and past the end of lt[] into ct[]); those reads must stay inside this
 * warp's record, hence ctl[] first. */
struct K2WarpSmem {
    __align__(16) MvgMbCtl ctl[2 * K2_CTL_CHUNK];       /* control records, two chunks: record of macroblock mx at [mx & 63]      */
    MVG_CANARY(c0)
    __align__(16) int16_t  resid[K2_RING][384];         /* residual ring, filled by per-lane async copies a pair of macroblocks ahead */
    MVG_CANARY(c1)
    __align__(16) uint8_t  lt[MVG_LT_ROWS * MVG_LT_STRIDE];
    __align__(16) uint8_t  ct[2][MVG_CT_PLANE];
    MVG_CANARY(c2)
    __align__(16) uint8_t  n8[MVG_N8_BYTES];            /* Intra8x8 neighbour line: 32 words {p', f2, f3, -}; [MVG_N8_DC] = DC */
    MVG_CANARY(c3)
};

/* dynamic shared memory: warp records and the tap tables (2 KB aligned, see the kernel) */
#define K2_LUT_BYTES  ((sizeof(MvgLuts) + 127) / 128 * 128)
#define K2_SMEM_BYTES (sizeof(K2WarpSmem) * K2_WARPS + 2048 + K2_LUT_BYTES)

/* 16-byte asynchronous copy global -> shared (LDGSTS), completion by per-thread groups */
__device__ __forceinline__ void mvg_cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(mvg_smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void mvg_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void mvg_cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ uint2 mvg_ld_relaxed_u64(const uint2 *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return make_uint2((unsigned)v, (unsigned)(v >> 32));
}
/* two neighbouring halo words with one request; each 64-bit half is a relaxed access of its own */
__device__ __forceinline__ uint4 mvg_ld_relaxed_2u64(const uint4 *p)
{
    unsigned long long a, b;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    return make_uint4((unsigned)a, (unsigned)(a >> 32), (unsigned)b, (unsigned)(b >> 32));
}
__device__ __forceinline__ void mvg_st_relaxed_u64(uint2 *p, unsigned lo, unsigned hi)
{
    const unsigned long long v = (unsigned long long)lo | ((unsigned long long)hi << 32);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

/* byte sum of a 32-bit word */
__device__ __forceinline__ int mvg_sum4(unsigned w) { return (int)__dp4a(w, 0x01010101u, 0u); }
/* Clip1(pred + r): one VIADDMNMX */
__device__ __forceinline__ int mvg_add_clip8(int pred, int r) { return __viaddmin_s32_relu(pred, r, 255); }
/* Clip1(pred + r) on two samples held as int16 pairs: one VIADDMNMX.S16x2.RELU.  The 16-bit add wraps,
 * which is why kernel 1 clamps the residual to [-512, 511]. */
__device__ __forceinline__ unsigned mvg_add_clip8x2(unsigned pred2, unsigned r2) { return __viaddmin_s16x2_relu(pred2, r2, 0x00ff00ffu); }
/* bytes 0,1 / 2,3 of a word as int16 pairs */
__device__ __forceinline__ unsigned mvg_pair_lo(unsigned w) { return __byte_perm(w, 0, 0x4140); }
__device__ __forceinline__ unsigned mvg_pair_hi(unsigned w) { return __byte_perm(w, 0, 0x4342); }
/* low bytes of two int16 pairs -> 4 bytes */
__device__ __forceinline__ unsigned mvg_pairs_to_bytes(unsigned a, unsigned b) { return __byte_perm(a, b, 0x6420); }

/* everything a lane needs to know about its warp's shared memory and its own role; the pointers are
 * warp-uniform (they come from a lane-0 broadcast, so they live in uniform registers and shared-memory
 * accesses take the form [lane register + uniform base + immediate]) */
struct K2Ctx {
#ifdef MVG_CHECKED
    const uint8_t *rec_lo, *rec_hi;     /* this warp's shared-memory record */
#endif
    uint8_t *lt, *ct;           /* luma tile, chroma tiles (plane stride MVG_CT_PLANE) */
    uint8_t *n8;
    const uint8_t *resid;       /* residual buffer of the current macroblock */
    const uint8_t *lut8;        /* MvgLuts::lut8[0][lane] in shared memory */
    int lane;
    /* Intra4x4: lane = 16 * half + 4 * py + px */
    unsigned lut4;              /* shared-memory address of lut4[0][lane] (table 2 KB aligned) */
    int s4;                     /* tile offset of my sample relative to the origin of the half-1 block */
    int r4odd, r4even;          /* residual byte offset relative to the half-0 block, by0 odd / even */
    unsigned h4;                /* tile offset of my block relative to the half-1 block (0 or 4 rows down, 8 left) */
    unsigned m4c, m4b, m4cc;    /* nibbles (bit 0) whose block has no up-right neighbour: always / if !availB / if !availC */
    const unsigned *sel;        /* K2Params::sel */
    /* Intra8x8: lane n = entry n of the neighbour line; its two samples: see MvgLuts::lut8 */
    int n8tr, n8notr;           /* tile offset of my neighbour sample relative to the block origin */
    int s8;
    unsigned fixA, fixB, fixD;  /* bit 2b: next := raw, bit 2b+1: prev := raw in block b when A / B / D is unavailable */
};

/* ---- Intra16x16 luma (h264_intra_prediction.c:1945-2141) ------------------- */
/* lane = 4 * p + row: row `row` of the horizontally adjacent 4x4 blocks 2p, 2p+1 (decoding order), i.e. eight
 * samples whose residual is two 8-byte reads at lane-linear addresses; handled as four int16 pairs */
__device__ __forceinline__ void k2_luma16(const K2Ctx &c, int mode, bool left, bool up)
{
    uint8_t *lt = c.lt;
    const int lane = c.lane, p = lane >> 2;
    const int x0 = (p & 2) * 4, y = ((p & 1) + (p >> 2) * 2) * 4 + (lane & 3);
    unsigned pp[4];
    if (mode == 0) {            /* Vertical */
        const uint2 t = *reinterpret_cast<const uint2 *>(lt + K2_TO(0, -1) + x0);
        pp[0] = mvg_pair_lo(t.x); pp[1] = mvg_pair_hi(t.x); pp[2] = mvg_pair_lo(t.y); pp[3] = mvg_pair_hi(t.y);
    } else if (mode == 1) {     /* Horizontal */
        pp[0] = pp[1] = pp[2] = pp[3] = (unsigned)lt[K2_TO(-1, 0) + y * MVG_LT_STRIDE] * 0x10001u;
    } else if (mode == 2) {     /* DC */
        int v = 0;
        if (lane < 16) { if (left) v = lt[K2_TO(-1, 0) + lane * MVG_LT_STRIDE]; }
        else if (lane < 20) { if (up) v = mvg_sum4(*reinterpret_cast<const unsigned *>(lt + K2_TO(0, -1) + (lane - 16) * 4)); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MVG_FULL, v, o);
        v = (left && up) ? (v + 16) >> 5 : (left || up) ? (v + 8) >> 4 : 128;
        pp[0] = pp[1] = pp[2] = pp[3] = (unsigned)v * 0x10001u;
    } else {                    /* Plane */
        int term = 0;
        const int i = lane & 7;
        if (lane < 8)       term = (i + 1) * ((int)lt[K2_TO(8 + i, -1)] - (int)lt[K2_TO(6 - i, -1)]);
        else if (lane < 16) term = (i + 1) * ((int)lt[K2_TO(-1, 8 + i)] - (int)lt[K2_TO(-1, 6 - i)]);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) term += __shfl_xor_sync(MVG_FULL, term, o);
        const int H = __shfl_sync(MVG_FULL, term, 0), V = __shfl_sync(MVG_FULL, term, 8);
        const int a = 16 * ((int)lt[K2_TO(-1, 15)] + (int)lt[K2_TO(15, -1)]);
        const int b = (5 * H + 32) >> 6, cc = (5 * V + 32) >> 6;
        /* v(x) = a + b (x - 7) + c (y - 7) + 16 fits int16 (|v| < 20000); Clip1(v >> 5) = clamp(v, 0, 8191) >> 5 */
        const int v0 = a + cc * (y - 7) + 16 + b * (x0 - 7);
        unsigned pair = __vadd2((unsigned)(v0 & 0xffff) * 0x10001u, (unsigned)b << 16);
        const unsigned step = (unsigned)((2 * b) & 0xffff) * 0x10001u;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            pp[k] = (__vimin_s16x2_relu(pair, 0x1fff1fffu) >> 5) & 0x00ff00ffu;
            pair = __vadd2(pair, step);
        }
    }
    const uint8_t *r = c.resid + (lane >> 2) * 64 + (lane & 3) * 8;
    const uint2 ra = *reinterpret_cast<const uint2 *>(r);
    const uint2 rb = *reinterpret_cast<const uint2 *>(r + 32);
    const unsigned lo = mvg_pairs_to_bytes(mvg_add_clip8x2(pp[0], ra.x), mvg_add_clip8x2(pp[1], ra.y));
    const unsigned hi = mvg_pairs_to_bytes(mvg_add_clip8x2(pp[2], rb.x), mvg_add_clip8x2(pp[3], rb.y));
    *reinterpret_cast<uint2 *>(lt + K2_TO(0, 0) + y * MVG_LT_STRIDE + x0) = make_uint2(lo, hi);
}

/* ---- Intra4x4 luma: anti-diagonal schedule, two blocks per step ------------- */
/* Blocks with bx + 2*by == t can be predicted together: their left, up, up-left and up-right neighbours
 * all belong to earlier steps.  Lanes 0..15 take block (t&1, t>>1), lanes 16..31 block ((t&1)+2, (t>>1)-1),
 * one sample per lane.  Every directional predictor is (n[a]+n[b]+n[c]+n[d]+2)>>2 over four (repeated)
 * neighbour samples; lut4[mode][lane] packs their four tile offsets (one byte each, relative to the
 * lane's block), so the table costs ONE shared-memory wavefront per step -- the shared-memory data pipe,
 * not instruction issue, bounds this kernel -- and a dot-product instruction per tap turns byte k into
 * an address: offset = dp4a(entry, 1 << 8k, lane's block offset).  Modes 3 and 7 without an up-right
 * neighbour use rows 11 and 15, whose taps stop at p[3,-1] (h264_intra_prediction.c:431-439). */
template <int T>
__device__ __forceinline__ void k2_luma4_step(const K2Ctx &c, unsigned seq, unsigned dcsteps)
{
    constexpr int bx0 = T & 1, by0 = T >> 1;
    constexpr bool v0 = T <= 7, v1 = T >= 2;
    constexpr int org1 = K2_TO(bx0 * 4 + 8, by0 * 4 - 4);       /* origin of the half-1 block */
    constexpr int blk0 = (bx0 & 1) | ((by0 & 1) << 1) | ((bx0 >> 1) << 2) | ((by0 >> 1) << 3);
    constexpr int sh = 4 * (T & 7);
    /* Every lane runs the whole step; in steps 0, 1, 8, 9 one half has no block and computes on whatever lies at
     * its (in-bounds) addresses -- only the store is conditional.  That keeps the shuffle in converged code. */
    const unsigned m = (sh >= 7 ? (seq >> (sh >= 7 ? sh - 7 : 0)) : (seq << (sh >= 7 ? 0 : 7 - sh))) & 0x780u;
    unsigned e;
    asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(m | c.lut4));
    const uint8_t *nb = c.lt + (org1 - MVG_LUT4_BIAS);
#ifdef MVG_CHECKED
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint8_t *a = nb + __dp4a(e, c.sel[k], c.h4);
        MVG_ASSERT(a >= c.rec_lo && a < c.rec_hi, 0);
    }
#endif
    int sum = (int)nb[__dp4a(e, c.sel[0], c.h4)] + (int)nb[__dp4a(e, c.sel[1], c.h4)] +
              (int)nb[__dp4a(e, c.sel[2], c.h4)] + (int)nb[__dp4a(e, c.sel[3], c.h4)] + 2;
#ifdef MVG_CHECKED
    {
        const uint8_t *a = c.resid + blk0 * 32 + ((by0 & 1) ? c.r4odd : c.r4even);
        MVG_ASSERT(a >= c.rec_lo && a + 2 <= c.rec_hi, 0);
    }
#endif
    int shift = 2;
    if (dcsteps & (1u << T)) {              /* warp-uniform: some block of this step is DC with both sides available */
        const int other = __shfl_xor_sync(MVG_FULL, sum, 1);
        if (m == 0x100u) { sum += other; shift = 3; }
    }
    const int r = *reinterpret_cast<const int16_t *>(c.resid + blk0 * 32 + ((by0 & 1) ? c.r4odd : c.r4even));
    const int v = mvg_add_clip8(sum >> shift, r);
    if ((v0 && v1) || (c.lane >= 16 ? v1 : v0)) c.lt[org1 + c.s4] = (uint8_t)v;
    __syncwarp();
}

/* nibbles equal to 2 (DC), as bit 0 of the nibble */
__device__ __forceinline__ unsigned k2_dc_nibbles(unsigned seq)
{
    return (seq >> 1) & ~seq & ~(seq >> 2) & ~(seq >> 3) & 0x11111111u;
}

__device__ __forceinline__ void k2_luma4(const K2Ctx &c, unsigned w1, unsigned w2, bool availA, bool availB, bool availC)
{
    const bool half = c.lane >= 16;
    unsigned seq = half ? w2 : w1;
    const unsigned notr = c.m4c | (availB ? 0u : c.m4b) | (availC ? 0u : c.m4cc);
    seq |= (seq & (seq >> 1) & notr) << 3;          /* modes 3, 7 -> 11, 15 where p[4..7,-1] are not available */
    if (!availA || !availB) {
        /* DC at the picture edge (h264_intra_prediction.c:554-600): blocks of the top row without a macroblock above
         * use the left samples only (row 10), blocks of the left column without a macroblock to the left the ones
         * above (row 9).  Block 0 with neither: the caller has set both neighbour lines to 128, it stays row 2. */
        const unsigned d = k2_dc_nibbles(seq);
        unsigned top = 0, left = 0;
        if (!availB) top = half ? 0x00001100u : (availA ? 0x00000011u : 0x00000010u);
        if (!availA && !half) left = availB ? 0x01010101u : 0x01010100u;
        seq |= (d & top) << 3;                      /* 2 -> 10 */
        seq ^= (d & left) * 0xBu;                   /* 2 -> 9  */
        if (!availA && !availB) {                   /* first macroblock of a picture: p[-1,0..3] = p[0..3,-1] = 128 */
            if (c.lane < 4) c.lt[K2_TO(-1, 0) + c.lane * MVG_LT_STRIDE] = 128;
            if (c.lane == 4) *reinterpret_cast<unsigned *>(c.lt + K2_TO(0, -1)) = 0x80808080u;
            __syncwarp();
        }
    }
    /* steps in which a block is DC with both sides: nibble t of either half -> bit t; half 0 covers steps 0..7 with
     * nibbles 0..7, half 1 steps 2..9 with nibbles 2..7, 0, 1 (see MvgMbCtl) */
    unsigned dcsteps;
    {
        const unsigned d = k2_dc_nibbles(seq);
        /* bit 4t -> bit t: pair the nibbles of a byte, then gather the four 2-bit fields with a multiplication
         * (fields land at bits 24, 26, 28, 30 of the product; all partial products are disjoint, so no carries) */
        unsigned bits = (((d | (d >> 3)) & 0x03030303u) * 0x01041040u) >> 24;
        if (half) bits = (bits & 0xfcu) | ((bits & 3u) << 8);
        dcsteps = __reduce_or_sync(MVG_FULL, bits);
    }
    k2_luma4_step<0>(c, seq, dcsteps);
    k2_luma4_step<1>(c, seq, dcsteps);
    k2_luma4_step<2>(c, seq, dcsteps);
    k2_luma4_step<3>(c, seq, dcsteps);
    k2_luma4_step<4>(c, seq, dcsteps);
    k2_luma4_step<5>(c, seq, dcsteps);
    k2_luma4_step<6>(c, seq, dcsteps);
    k2_luma4_step<7>(c, seq, dcsteps);
    k2_luma4_step<8>(c, seq, dcsteps);
    k2_luma4_step<9>(c, seq, dcsteps);
}

/* ---- Intra8x8 luma: 4 blocks in order ------------------------------------------ */
/* The 25 neighbours of a block form one line: n = 0..7 p[-1,7..0], 8 p[-1,-1], 9..24 p[0..15,-1].
 * Lane n filters its entry (reference sample filter, h264_intra_prediction.c:1295-1353) and then
 * derives, again with shuffles, the two smoothings every directional mode is built from:
 *   f2[n] = (p'[n] + p'[n+1] + 1) >> 1,  f3[n] = (p'[n-1] + 2 p'[n] + p'[n+1] + 2) >> 2
 * (line ends replicate).  Each predicted sample is then ONE of p'[i], f2[i], f3[i]; the table gives,
 * for the two horizontally adjacent samples of a lane, the word to load and the bit shift of the byte. */
template <int B8>
__device__ __forceinline__ void k2_luma8_block(const K2Ctx &c, unsigned modes, unsigned fix, bool availA, bool availB, bool availC)
{
    uint8_t *lt = c.lt;
    const int lane = c.lane;
    constexpr int xo = (B8 & 1) * 8, yo = (B8 >> 1) * 8;
    constexpr int org = K2_TO(xo, yo);
    const unsigned mode = (modes >> (4 * B8)) & 15u;
    const bool left = (B8 & 1) ? true : availA, up = (B8 & 2) ? true : availB;
    const bool tr = B8 == 0 ? availB : (B8 == 1 ? availC : (B8 == 2));

    const int raw = lt[org + (tr ? c.n8tr : c.n8notr)];
    int prev = __shfl_up_sync(MVG_FULL, raw, 1), next = __shfl_down_sync(MVG_FULL, raw, 1);
    /* lane 0 gets its own value back (p'[-1,7] = (p[-1,6] + 3 p[-1,7] + 2) >> 2); lane 25 loads what lane 24
     * loads (p'[15,-1]); the corner and its neighbours replicate when a side is missing */
    if (B8 != 3) {
        if (fix & (1u << (2 * B8))) next = raw;
        if (fix & (2u << (2 * B8))) prev = raw;
    }
    const int filt = (prev + 2 * raw + next + 2) >> 2;
    const int fp = __shfl_up_sync(MVG_FULL, filt, 1);
    int fn = __shfl_down_sync(MVG_FULL, filt, 1);
    if (lane == 24) fn = filt;
    const int f2 = (filt + fn + 1) >> 1, f3 = (fp + 2 * filt + fn + 2) >> 2;
    /* p', f2, f3 of my line entry as ONE word (all three are 0..255): one shared-memory wavefront instead of three */
    reinterpret_cast<unsigned *>(c.n8)[lane] = (unsigned)filt | ((unsigned)f2 << 8) | ((unsigned)f3 << 16);
    if (mode == 2) {                                        /* warp-uniform */
        int v = 0;
        if (lane < 8 && left) v = filt;
        if (lane >= 9 && lane < 17 && up) v = filt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MVG_FULL, v, o);
        v = (left && up) ? (v + 8) >> 4 : (left || up) ? (v + 4) >> 3 : 128;
        if (lane == 0) c.n8[MVG_N8_DC] = (uint8_t)v;
    }
    __syncwarp();
    const unsigned e = *reinterpret_cast<const unsigned *>(c.lut8 + mode * 128);
    MVG_ASSERT(mode < 16 && (e & 0xffffu) < MVG_N8_BYTES && (e >> 16) < MVG_N8_BYTES, 3);
    const unsigned pp = (unsigned)c.n8[e & 0xffffu] | ((unsigned)c.n8[e >> 16] << 16);
    const unsigned r2 = *reinterpret_cast<const unsigned *>(c.resid + B8 * 128 + lane * 4);
    *reinterpret_cast<uint16_t *>(lt + org + c.s8) = (uint16_t)__byte_perm(mvg_add_clip8x2(pp, r2), 0, 0x4420);
    __syncwarp();
}

/* the same with the block index at run time: one copy of the code instead of four (the fused kernel has to fit
 * the instruction cache), a handful of address instructions more per block */
__device__ __forceinline__ void k2_luma8_block_rt(const K2Ctx &c, int b8, unsigned modes, unsigned fix, bool availA, bool availB, bool availC)
{
    uint8_t *lt = c.lt;
    const int lane = c.lane;
    const int org = K2_TO(0, 0) + (b8 & 1) * 8 + (b8 >> 1) * 8 * MVG_LT_STRIDE;
    const unsigned mode = (modes >> (4 * b8)) & 15u;
    const bool left = (b8 & 1) ? true : availA, up = (b8 & 2) ? true : availB;
    const bool tr = b8 == 0 ? availB : (b8 == 1 ? availC : (b8 == 2));

    const int raw = lt[org + (tr ? c.n8tr : c.n8notr)];
    int prev = __shfl_up_sync(MVG_FULL, raw, 1), next = __shfl_down_sync(MVG_FULL, raw, 1);
    {
        const unsigned f2 = b8 == 3 ? 0u : fix >> (2 * b8);
        if (f2 & 1u) next = raw;
        if (f2 & 2u) prev = raw;
    }
    const int filt = (prev + 2 * raw + next + 2) >> 2;
    const int fp = __shfl_up_sync(MVG_FULL, filt, 1);
    int fn = __shfl_down_sync(MVG_FULL, filt, 1);
    if (lane == 24) fn = filt;
    const int f2 = (filt + fn + 1) >> 1, f3 = (fp + 2 * filt + fn + 2) >> 2;
    /* p', f2, f3 of my line entry as ONE word (all three are 0..255): one shared-memory wavefront instead of three */
    reinterpret_cast<unsigned *>(c.n8)[lane] = (unsigned)filt | ((unsigned)f2 << 8) | ((unsigned)f3 << 16);
    if (mode == 2) {                                        /* warp-uniform */
        int v = 0;
        if (lane < 8 && left) v = filt;
        if (lane >= 9 && lane < 17 && up) v = filt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MVG_FULL, v, o);
        v = (left && up) ? (v + 8) >> 4 : (left || up) ? (v + 4) >> 3 : 128;
        if (lane == 0) c.n8[MVG_N8_DC] = (uint8_t)v;
    }
    __syncwarp();
    const unsigned e = *reinterpret_cast<const unsigned *>(c.lut8 + mode * 128);
    const unsigned pp = (unsigned)c.n8[e & 0xffffu] | ((unsigned)c.n8[e >> 16] << 16);
    const unsigned r2 = *reinterpret_cast<const unsigned *>(c.resid + b8 * 128 + lane * 4);
    *reinterpret_cast<uint16_t *>(lt + org + c.s8) = (uint16_t)__byte_perm(mvg_add_clip8x2(pp, r2), 0, 0x4420);
    __syncwarp();
}

template <bool COMPACT = false>
__device__ __forceinline__ void k2_luma8(const K2Ctx &c, unsigned modes, bool availA, bool availB, bool availC, bool availD)
{
    const unsigned fix = (availA ? 0u : c.fixA) | (availB ? 0u : c.fixB) | (availD ? 0u : c.fixD);
    if (COMPACT) {
#pragma unroll 1
        for (int b8 = 0; b8 < 4; b8++) k2_luma8_block_rt(c, b8, modes, fix, availA, availB, availC);
        return;
    }
    k2_luma8_block<0>(c, modes, fix, availA, availB, availC);
    k2_luma8_block<1>(c, modes, fix, availA, availB, availC);
    k2_luma8_block<2>(c, modes, fix, availA, availB, availC);
    k2_luma8_block<3>(c, modes, fix, availA, availB, availC);
}

/* ---- chroma, both planes at once (h264_intra_prediction.c:2338-2564) --------- */
/* lane = 16 * plane + 4 * block + row: four samples of one row of one 4x4 block as two int16 pairs; the
 * residual of lane l is the 8 bytes at 512 + 8 l */
__device__ __forceinline__ void k2_chroma(const K2Ctx &c, int mode, bool left, bool up)
{
    const int lane = c.lane, pl = lane >> 4;
    const int x0 = (lane & 4), yo = (lane & 8) >> 1, y = yo + (lane & 3);
    uint8_t *ct = c.ct + pl * MVG_CT_PLANE;
    unsigned p0, p1;
    if (mode == 0) {            /* DC, per 4x4 block */
        int st = 0, sl = 0;
        if (up) st = mvg_sum4(*reinterpret_cast<const unsigned *>(ct + K2_CO(0, -1) + x0));
        if (left) sl = (int)ct[K2_CO(-1, 0) + yo * MVG_CT_STRIDE] + (int)ct[K2_CO(-1, 1) + yo * MVG_CT_STRIDE] +
                       (int)ct[K2_CO(-1, 2) + yo * MVG_CT_STRIDE] + (int)ct[K2_CO(-1, 3) + yo * MVG_CT_STRIDE];
        int v;
        if (!left && !up) v = 128;
        else if ((x0 == 0) == (yo == 0))       /* blocks (0,0) and (4,4) */
            v = (left && up) ? (st + sl + 4) >> 3 : left ? (sl + 2) >> 2 : (st + 2) >> 2;
        else if (x0 > 0) v = up ? (st + 2) >> 2 : (sl + 2) >> 2;          /* (4,0): top first  */
        else v = left ? (sl + 2) >> 2 : (st + 2) >> 2;                    /* (0,4): left first */
        p0 = p1 = (unsigned)v * 0x10001u;
    } else if (mode == 1) {     /* Horizontal */
        p0 = p1 = (unsigned)ct[K2_CO(-1, 0) + y * MVG_CT_STRIDE] * 0x10001u;
    } else if (mode == 2) {     /* Vertical */
        const unsigned t = *reinterpret_cast<const unsigned *>(ct + K2_CO(0, -1) + x0);
        p0 = mvg_pair_lo(t); p1 = mvg_pair_hi(t);
    } else {                    /* Plane */
        /* the eight products of a plane, one per lane (lanes 0..3 of a plane: H, 4..7: V, the rest contribute 0),
         * summed with two butterfly shuffles and handed to all 16 lanes of the plane with two more */
        const int i = lane & 3, g = (lane >> 2) & 3;
        const int step = g == 0 ? 1 : MVG_CT_STRIDE;                                    /* along the top row / down the left column */
        const int mid = g == 0 ? K2_CO(3, -1) : K2_CO(-1, 3);                           /* p[3,-1] resp. p[-1,3] */
        const int term = (g < 2 ? i + 1 : 0) * ((int)ct[mid + (i + 1) * step] - (int)ct[mid - (i + 1) * step]);
        int sum = term + __shfl_xor_sync(MVG_FULL, term, 1);
        sum += __shfl_xor_sync(MVG_FULL, sum, 2);
        const int H = __shfl_sync(MVG_FULL, sum, lane & 16), V = __shfl_sync(MVG_FULL, sum, (lane & 16) + 4);
        const int a = 16 * ((int)ct[K2_CO(-1, 7)] + (int)ct[K2_CO(7, -1)]);
        const int b = (34 * H + 32) >> 6, cc = (34 * V + 32) >> 6;
        /* |a + b (x - 3) + c (y - 3) + 16| < 2^15: int16 pairs as in the luma plane predictor */
        const int v0 = a + cc * (y - 3) + 16 + b * (x0 - 3);
        const unsigned pair = __vadd2((unsigned)(v0 & 0xffff) * 0x10001u, (unsigned)b << 16);
        p0 = (__vimin_s16x2_relu(pair, 0x1fff1fffu) >> 5) & 0x00ff00ffu;
        p1 = (__vimin_s16x2_relu(__vadd2(pair, (unsigned)((2 * b) & 0xffff) * 0x10001u), 0x1fff1fffu) >> 5) & 0x00ff00ffu;
    }
    const uint2 r = *reinterpret_cast<const uint2 *>(c.resid + 512 + lane * 8);
    *reinterpret_cast<unsigned *>(ct + K2_CO(0, 0) + y * MVG_CT_STRIDE + x0) =
        mvg_pairs_to_bytes(mvg_add_clip8x2(p0, r.x), mvg_add_clip8x2(p1, r.y));
}

/* Persistent warps.  A work item is one macroblock row of one picture; a warp claims items
 * from an atomic counter and walks its row left to right.
 *
 * Wavefront dependency: row r may process macroblock x once row r-1 has finished macroblock
 * x+1 (its up-right neighbour C, h264_spatial.c:371-382).  The only samples that cross rows are
 * the bottom sample line of the macroblocks above, so a finished macroblock publishes that
 * line (16 Y + 8 Cb + 8 Cr bytes) as eight 64-bit words {4 data bytes, launch epoch} -- the
 * flag travels with the data, every word validates itself, and neither side needs a fence or
 * a separate progress counter.  The row below reads them four macroblocks at a time (one coalesced
 * load, a group ahead of use) and polls only when a word it needs is stale.
 *
 * Claim order: pictures are taken in groups of `group`; inside a group items are ordered
 * row-major over (row, picture).  Row r-1 of a picture is therefore always claimed before row
 * r (no deadlock: it runs on a resident warp).
 *
 * Inputs arrive by per-lane asynchronous copies (LDGSTS, completion by per-thread groups, no barrier
 * objects): the residuals of macroblocks x+2 and x+3 (1536 bytes: three full-warp 16-byte copies, no lane
 * condition) are requested while x is predicted, control records come 32 macroblocks at a time.  Output is one 384-byte tile per macroblock.
 *
 * What bounds the kernel (ncu, profiles/): instruction issue and the shared-memory data pipe, not HBM;
 * hence tables with ready-to-use offsets, [lane constant + uniform base + immediate] addressing,
 * int16-pair residual adds, and lane mappings that make addresses linear in the lane index. */
__global__ void __launch_bounds__(K2_WARPS * 32, 2)
k2_wavefront(K2Params p)
{
    extern __shared__ __align__(128) uint8_t k2_smem[];
    const int lane = mvg_lane();
    const unsigned wid = __shfl_sync(MVG_FULL, threadIdx.x >> 5, 0);       /* warp-uniform by construction */
    /* layout: the tap tables sit on the first 2 KB boundary (so that (mode << 7) can be OR-ed into a lane's
     * table address); warp records fill the space before it, the others follow the tables */
    const unsigned base = mvg_smem_u32(k2_smem);
    const unsigned lut_addr = (base + 2047u) & ~2047u;
    const unsigned n_before = (lut_addr - base) / (unsigned)sizeof(K2WarpSmem);
    MvgLuts *luts = reinterpret_cast<MvgLuts *>(k2_smem + (lut_addr - base));
    K2WarpSmem &s = *reinterpret_cast<K2WarpSmem *>(
        wid < n_before ? k2_smem + wid * sizeof(K2WarpSmem)
                       : k2_smem + (lut_addr - base) + K2_LUT_BYTES + (wid - n_before) * sizeof(K2WarpSmem));
    for (int i = threadIdx.x; i < (int)(K2_LUT_BYTES / 16); i += blockDim.x)
        reinterpret_cast<uint4 *>(luts)[i] = __ldg(reinterpret_cast<const uint4 *>(p.luts) + i);
    __syncthreads();

    const int W = p.w_mbs, H = p.h_mbs, n_mb = W * H;
    const int total = p.n_pics * H;
    const unsigned epoch = p.epoch;

    /* per-lane constants --------------------------------------------------------------- */
    K2Ctx c;
#ifdef MVG_CHECKED
    c.rec_lo = reinterpret_cast<const uint8_t *>(&s); c.rec_hi = c.rec_lo + sizeof(K2WarpSmem);
    if (lane < 4) s.c0[lane] = s.c1[lane] = s.c2[lane] = s.c3[lane] = MVG_CANARY_WORD;
    __syncwarp();
#endif
    c.lt = s.lt; c.ct = &s.ct[0][0]; c.n8 = s.n8; c.lut8 = reinterpret_cast<const uint8_t *>(&luts->lut8[0][lane]);
    c.lane = lane;
    c.sel = p.sel;
    c.resid = reinterpret_cast<const uint8_t *>(s.resid[0]);
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
        /* Intra8x8 neighbour gather, relative to the block origin */
        const int n = lane < 25 ? lane : 24;
        if (n < 8)       c.n8tr = (7 - n) * MVG_LT_STRIDE - 1;
        else if (n == 8) c.n8tr = -MVG_LT_STRIDE - 1;
        else             c.n8tr = -MVG_LT_STRIDE + (n - 9);
        c.n8notr = n > 16 ? -MVG_LT_STRIDE + 7 : c.n8tr;                    /* p[8..15,-1] := p[7,-1] */
        /* Intra8x8 samples of a lane: 4x4 sub-block lane>>3, row (lane>>1)&3, pair lane&1, so that the residual of
         * the pair is the word at 4 * lane of the 8x8 block */
        const int x8 = ((lane >> 3) & 1) * 4 + (lane & 1) * 2, y8 = (lane >> 4) * 4 + ((lane >> 1) & 3);
        c.s8 = y8 * MVG_LT_STRIDE + x8;
        /* replicate at the corner: block 0 sees A, B, D of the macroblock; block 1: left = block 0, up and
         * up-left from B; block 2: up = block 0, left and up-left from A */
        c.fixA = lane == 7 ? 0x10u : lane == 8 ? 0x22u : lane == 9 ? 0x20u : 0u;
        c.fixB = lane == 7 ? 0x04u : lane == 8 ? 0x05u : lane == 9 ? 0x08u : 0u;
        c.fixD = lane == 7 ? 0x01u : lane == 9 ? 0x02u : 0u;
    }
    /* sample row -1 of the tiles comes from the halo words of the row above: lanes 0..3 luma x = 4*lane,
     * 4,5 Cb, 6,7 Cr of the macroblock above, lanes 8,9 luma x = 16..23 of the macroblock above-right */
    uint8_t *const halo_top = lane < 4 ? s.lt + K2_TO(lane * 4, -1)
                            : lane < 8 ? s.ct[(lane >> 1) & 1] + K2_CO((lane & 1) * 4, -1)
                                       : s.lt + K2_TO(16 + (lane & 1) * 4, -1);
    const uint8_t *const halo_bot = lane < 4 ? halo_top + 16 * MVG_LT_STRIDE : halo_top + 8 * MVG_CT_STRIDE;
    /* tile write-out: lanes 0..15 one luma row (16 B as two 8-byte pieces), lanes 16..23 Cb rows, 24..31 Cr rows (8 B) */
    const uint8_t *const wo_src = lane < 16 ? s.lt + K2_TO(0, lane)
                                            : s.ct[(lane >> 3) & 1] + K2_CO(0, lane & 7);
    const int wo_off = lane < 16 ? lane * 16 : 256 + (lane - 16) * 8;
    /* column x = 15 (luma) / x = 7 (chroma) -> x = -1 hand-over to the next macroblock: the last byte of what a
     * lane has just loaded for the write-out, stored one column left of its row ... */
    uint8_t *const lc_dst = lane < 16 ? s.lt + K2_TO(-1, lane) : s.ct[(lane >> 3) & 1] + K2_CO(-1, lane & 7);
    const unsigned lc_sel = lane < 16 ? 7u : 3u;            /* byte 3 of the second / first 8-byte piece */
    /* ... and row -1: lanes 3, 5, 7 hold the last word of the luma / Cb / Cr line above (see halo_top); the
     * other lanes copy an unused byte onto itself so that the move needs no predicate */
    const uint8_t *const cn_src = (lane == 3 || lane == 5 || lane == 7) ? halo_top + 3 : s.lt;
    uint8_t *const cn_dst = lane == 3 ? s.lt + K2_TO(-1, -1) : lane == 5 ? s.ct[0] + K2_CO(-1, -1)
                          : lane == 7 ? s.ct[1] + K2_CO(-1, -1) : s.lt;
    /* staging: lane l copies bytes [16 l, 16 l + 16) of a residual and, lanes 0..15, [512 + 16 l, ..) */
    uint8_t *const st_dst = reinterpret_cast<uint8_t *>(s.resid[0]) + lane * 16;

    K2_PROF(long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt0 = clock64();)

    for (;;) {
        K2_PROF(const long long tr0 = clock64();)
        int item = 0;
        if (lane == 0) item = atomicAdd(p.work, 1);
        item = __shfl_sync(MVG_FULL, item, 0);
        if (item >= total) break;
        const int g = item / (p.group * H);
        const int gsize = min(p.group, p.n_pics - g * p.group);
        const int within = item - g * p.group * H;
        const int row = within / gsize;
        const int slot = p.first_slot + g * p.group + (within - row * gsize);
        const size_t mb0 = (size_t)slot * n_mb + (size_t)row * W;           /* first macroblock of the row */

        const uint8_t *st_src = reinterpret_cast<const uint8_t *>(p.resid + mb0 * 384) + lane * 16;
        const MvgMbCtl *ctl = p.ctl + mb0;
        /* residuals travel in pairs of macroblocks: 1536 bytes are three full-warp 16-byte copies, no lane condition.
         * group 0: control records 0..31 and residuals 0, 1 */
        if (lane < W) mvg_cp_async16(&s.ctl[lane], ctl + lane);
        mvg_cp_async16(st_dst, st_src);
        if (W > 1) { mvg_cp_async16(st_dst + 512, st_src + 512); mvg_cp_async16(st_dst + 1024, st_src + 1024); }
        else if (lane < 16) mvg_cp_async16(st_dst + 512, st_src + 512);
        mvg_cp_async_commit();

        /* running pointers: source of the residual requested next (macroblock mx + 2), this lane's piece of the
         * tile of macroblock mx, its word of the published line, its word of the next group of the row above */
        const uint8_t *st_run = st_src + 2 * 768;
        uint8_t *wo_run = p.tiles + mb0 * 384 + wo_off;
        uint2 *hm_run = p.halo + mb0 * 8 + lane;
        const uint2 *ha_run = p.halo + (mb0 - W) * 8 + lane;           /* group of macroblock mx (at j == 0) */
        const bool availB = row > 0, publish = row < H - 1;
        const int hwords = W * 8;                               /* halo words of a macroblock row */
        uint2 qa = make_uint2(0, epoch), qb = make_uint2(0, epoch);
        if (availB) {
            if (lane < hwords) qb = mvg_ld_relaxed_u64(ha_run);     /* becomes qa at macroblock 0 */
        }
        unsigned okA = 0;

        K2_PROF(pc[0] += clock64() - tr0;)
        for (int c0 = 0; c0 < W; c0 += K2_CTL_CHUNK) {
            /* next chunk of control records; joins the copy group of the first macroblock of this chunk */
            if (c0 + K2_CTL_CHUNK + lane < W)
                mvg_cp_async16(&s.ctl[(c0 + K2_CTL_CHUNK + lane) & (2 * K2_CTL_CHUNK - 1)], ctl + c0 + K2_CTL_CHUNK + lane);
            const int cend = min(c0 + K2_CTL_CHUNK, W);
            for (int mx = c0; mx < cend; mx++) {
                K2_PROF(const long long t0 = clock64();)
                const int j = mx & 3;
                /* every second macroblock: request the pair mx + 2, mx + 3; their ring slots were last read before the
                 * __syncwarp() that closed mx - 1 */
                if ((mx & 1) == 0) {
                    const int left = W - (mx + 2);
                    uint8_t *d = st_dst + ((mx + 2) & (K2_RING - 1)) * 768;
                    if (left >= 2) {
                        mvg_cp_async16(d, st_run); mvg_cp_async16(d + 512, st_run + 512); mvg_cp_async16(d + 1024, st_run + 1024);
                    } else if (left == 1) {
                        mvg_cp_async16(d, st_run);
                        if (lane < 16) mvg_cp_async16(d + 512, st_run + 512);
                    }
                    st_run += 2 * 768;
                    mvg_cp_async_commit();
                }
                const bool availA = mx > 0, availC = availB && mx < W - 1, availD = availA && availB;

                if (availB) {
                    if (j == 0) {               /* group of four macroblocks above: requested a group ago */
                        qa = qb;
                        okA = __ballot_sync(MVG_FULL, qa.y == epoch);
                    }
                    /* words needed now: the 8 of the macroblock above and, for the up-right neighbour, the first
                     * two of the next one, which sit in qb when this is the last macroblock of the group */
                    const unsigned need = availC ? 0x3FFu : 0xFFu;
                    unsigned have = __funnelshift_r(okA, j == 3 ? __ballot_sync(MVG_FULL, qb.y == epoch) : 0u, 8 * j);
                    K2_PROF(const long long th = clock64();)
                    if ((have & need) != need) {
                        K2_PROF(pc[5]++;)
                        /* this row has caught up with the row above: poll, sleeping a fraction of a macroblock time */
                        unsigned ns = K2_POLL_NS;
                        do {
                            K2_PROF(pc[6]++;)
                            __nanosleep(ns);
                            if (ns < 8 * K2_POLL_NS) ns *= 2;
                            if ((mx & ~3) * 8 + lane < hwords) qa = mvg_ld_relaxed_u64(ha_run);
                            if ((mx & ~3) * 8 + 32 + lane < hwords) qb = mvg_ld_relaxed_u64(ha_run + 32);
                            okA = __ballot_sync(MVG_FULL, qa.y == epoch);
                            have = __funnelshift_r(okA, __ballot_sync(MVG_FULL, qb.y == epoch), 8 * j);
                        } while ((have & need) != need);
                    }
                    K2_PROF(pc[7] += clock64() - th;)
                    /* sample row -1 of the tiles: lanes 0..7 the macroblock above, lanes 8,9 x = 16..23 */
                    const unsigned src = (j == 3 && lane < 2) ? qb.x : qa.x;
                    const unsigned v = __shfl_sync(MVG_FULL, src, (8 * j + lane) & 31);
                    if (lane < 10) *reinterpret_cast<unsigned *>(halo_top) = v;
                    if (j == 3) ha_run += 32;
                }
                K2_PROF(const long long t1 = clock64();)
                mvg_cp_async_wait<1>();         /* all but the youngest group: the pair of macroblock mx has landed */
                __syncwarp();
                if (availB && j == 0) {
                    /* request the next group of the row above only now, behind everything that reads qb: those reads
                     * are predicated per macroblock, and a predicated-off read still waits for a load in flight.  The
                     * group is first needed three macroblocks from here. */
                    qb = make_uint2(0, epoch);
                    if (mx * 8 + 32 + lane < hwords) qb = mvg_ld_relaxed_u64(ha_run + 32);
                }
                K2_PROF(const long long t2 = clock64();)
                c.resid = reinterpret_cast<const uint8_t *>(s.resid[mx & (K2_RING - 1)]);
                const uint4 ctlw = *reinterpret_cast<const uint4 *>(&s.ctl[mx & (2 * K2_CTL_CHUNK - 1)]);

                const int kind = ctlw.x & 255, i16 = (ctlw.x >> 8) & 255, cmode = (ctlw.x >> 16) & 255;
                if (kind == MVG_MB_I16x16)    k2_luma16(c, i16, availA, availB);
                else if (kind == MVG_MB_I4x4) k2_luma4(c, ctlw.y, ctlw.z, availA, availB, availC);
                else                          k2_luma8(c, ctlw.y, availA, availB, availC, availD);
                k2_chroma(c, cmode, availA, availB);
                __syncwarp();
                K2_PROF(const long long t3 = clock64();)

                /* write the macroblock out as one 384-byte tile (coalesced; scattering 16-byte row pieces over a
                 * planar picture costs more than the whole prediction: measured 4.5 ms vs 1.8 ms per 1000 pictures) */
                const uint2 wa = *reinterpret_cast<const uint2 *>(wo_src);
                uint2 wb = make_uint2(0u, 0u);
                if (lane < 16) {
                    wb = *reinterpret_cast<const uint2 *>(wo_src + 8);
                    *reinterpret_cast<uint4 *>(wo_run) = make_uint4(wa.x, wa.y, wb.x, wb.y);
                } else *reinterpret_cast<uint2 *>(wo_run) = wa;
                wo_run += 384;
                /* publish the bottom sample line for the row below */
                if (publish && lane < 8)
                    mvg_st_relaxed_u64(hm_run, *reinterpret_cast<const unsigned *>(halo_bot), epoch);
                hm_run += 8;
                /* next macroblock: x = 15 becomes x = -1 (luma rows -1..15, chroma x = 7, rows -1..7) */
                *lc_dst = (uint8_t)__byte_perm(wa.y, wb.y, lc_sel);
                *cn_dst = *cn_src;
                __syncwarp();
                K2_PROF(const long long t4 = clock64(); pc[1] += t1 - t0; pc[2] += t2 - t1; pc[3] += t3 - t2; pc[4] += t4 - t3;)
            }
        }
        mvg_cp_async_wait<0>();
    }
    K2_PROF(if (lane == 0 && p.stats) { for (int i = 0; i < 8; i++) atomicAdd(p.stats + i, (unsigned long long)pc[i]); atomicAdd(p.stats + 8, (unsigned long long)(clock64() - pt0)); })
#ifdef MVG_CHECKED
    __syncwarp();
    if (lane < 4) MVG_ASSERT(s.c0[lane] == MVG_CANARY_WORD && s.c1[lane] == MVG_CANARY_WORD && s.c2[lane] == MVG_CANARY_WORD && s.c3[lane] == MVG_CANARY_WORD, 2);
#endif
}

/* ========================================================================= */
/* Kernel 3                                                                    */

struct K3Params {
    const uint8_t *tiles;   /* [slot][n_mb][384] */
    uint8_t       *rgb;     /* [slot][3*(W/s)*(H/s)] */
    uint8_t       *yuv;     /* [slot][1.5*W*H] planar I420 (k4 only) */
    int width, height, scale, first_slot, n_pics;
};

/* export_utils.c:300-302 */
__device__ __forceinline__ void mvg_ycc_to_rgb(int Y, int Cb, int Cr, int &R, int &G, int &B)
{
    const int t = (298 * Y) >> 8;
    R = mvg_clip8(t + ((408 * Cr) >> 8) - 222);
    G = mvg_clip8(t - ((100 * Cb) >> 8) - ((208 * Cr) >> 8) + 135);
    B = mvg_clip8(t + ((516 * Cb) >> 8) - 276);
}

/* int16 pairs {a.b[lo], a.b[hi]} of a byte word */
__device__ __forceinline__ unsigned mvg_pair_even(unsigned w) { return __byte_perm(w, 0, 0x4240); }   /* bytes 0, 2 */
__device__ __forceinline__ unsigned mvg_pair_odd(unsigned w)  { return __byte_perm(w, 0, 0x4341); }   /* bytes 1, 3 */

/* scale 1, byte for byte mb_to_rgb() (export_utils.c:266-303), which walks the reference's per-macroblock
 * sample arrays exactly as this kernel walks the tiles.  A CTA of 8 warps converts 32 consecutive macroblocks:
 * their tiles (12 KB, contiguous in HBM) are copied to shared memory with coalesced 128-bit loads, then warp
 * q takes luma rows 2q, 2q+1 of all 32 macroblocks, lane = macroblock: 2 x 16 luma samples + 8 Cb + 8 Cr in,
 * 2 x 48 bytes of RGB24 out, and the 32 lanes of a warp write 1536 contiguous bytes per picture row.  (Letting
 * every lane fetch its row pair straight from its tile, 384 bytes from its neighbour's, costs 30 % more time;
 * four macroblocks per warp with coalesced loads scatters the stores over 16 picture rows and costs 50 %.)
 * The tile stride in shared memory is 400 bytes: 16-byte reads at that stride are bank-conflict free.
 * The arithmetic runs on int16 pairs -- pixels x and x+2 of a row, which use chroma samples c and c+1:
 *   (298 Y) >> 8 = (149 Y) >> 7,  (408 Cr) >> 8 = (204 Cr) >> 7,  (516 Cb) >> 8 = (129 Cb) >> 6,
 *   (100 Cb) >> 8 = (25 Cb) >> 6,  (208 Cr) >> 8 = (13 Cr) >> 4          (all products < 2^16: no carry between
 * the halves of a packed multiply), the clip is one VIADDMNMX.S16x2.RELU per colour and pixel pair. */
#define K3_MBS   32          /* macroblocks per CTA iteration */
#define K3_TILE  400         /* tile stride in shared memory  */
__global__ void __launch_bounds__(256)
k3_rgb_full(K3Params p)
{
    __shared__ __align__(16) uint8_t s_tiles[K3_MBS * K3_TILE];
    const int w_mbs = p.width >> 4, h_mbs = p.height >> 4, n_mb = w_mbs * h_mbs;
    const long long total_mbs = (long long)n_mb * p.n_pics;
    const long long n_groups = (total_mbs + K3_MBS - 1) / K3_MBS;
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
    const size_t ysz = (size_t)p.width * p.height;
    const uint8_t *tiles = p.tiles + (size_t)p.first_slot * n_mb * 384;
    uint8_t *rgb = p.rgb + (size_t)p.first_slot * ysz * 3;
    /* position of this lane's macroblock, advanced by the grid stride without divisions */
    long long mb = (long long)blockIdx.x * K3_MBS + lane;
    int pic = (int)(mb / n_mb), my = (int)((mb - (long long)pic * n_mb) / w_mbs), mx = (int)(mb - (long long)pic * n_mb - (long long)my * w_mbs);
    const long long stride = (long long)gridDim.x * K3_MBS;
    const int dpic = (int)(stride / n_mb), dmy = (int)((stride - (long long)dpic * n_mb) / w_mbs),
              dmx = (int)(stride - (long long)dpic * n_mb - (long long)dmy * w_mbs);
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x, mb += stride) {
        {   /* 32 tiles = 768 chunks of 16 bytes, three per thread */
            const long long first = g * K3_MBS;
            const int n_chunks = (int)min((long long)K3_MBS, total_mbs - first) * 24;
            const uint4 *src = reinterpret_cast<const uint4 *>(tiles + (size_t)first * 384);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int ch = threadIdx.x + 256 * k;
                if (ch < n_chunks) {
                    const int t = ch / 24;
                    *reinterpret_cast<uint4 *>(s_tiles + t * K3_TILE + (ch - t * 24) * 16) = __ldg(src + ch);
                }
            }
        }
        __syncthreads();
        if (mb < total_mbs) {
            const uint8_t *tile = s_tiles + lane * K3_TILE;
            const uint4 y0 = *reinterpret_cast<const uint4 *>(tile + q * 32);
            const uint4 y1 = *reinterpret_cast<const uint4 *>(tile + q * 32 + 16);
            const uint2 cb = *reinterpret_cast<const uint2 *>(tile + 256 + q * 8);
            const uint2 cr = *reinterpret_cast<const uint2 *>(tile + 320 + q * 8);
            const unsigned yw[2][4] = {{y0.x, y0.y, y0.z, y0.w}, {y1.x, y1.y, y1.z, y1.w}};
            const unsigned cbw[2] = {cb.x, cb.y}, crw[2] = {cr.x, cr.y};
            unsigned out[2][12];
#pragma unroll
            for (int k = 0; k < 4; k++) {               /* luma samples 4k..4k+3 of both rows, chroma samples 2k, 2k+1 */
                const unsigned cb2 = (k & 1) ? mvg_pair_hi(cbw[k >> 1]) : mvg_pair_lo(cbw[k >> 1]);
                const unsigned cr2 = (k & 1) ? mvg_pair_hi(crw[k >> 1]) : mvg_pair_lo(crw[k >> 1]);
                /* export_utils.c:300-302, the terms that do not depend on Y */
                const unsigned rC = __vsub2(((cr2 * 204u) >> 7) & 0x01ff01ffu, 0x00de00deu);                   /* - 222 */
                const unsigned bC = __vsub2(((cb2 * 129u) >> 6) & 0x03ff03ffu, 0x01140114u);                   /* - 276 */
                const unsigned gC = __vsub2(__vsub2(0x00870087u, ((cb2 * 25u) >> 6) & 0x00ff00ffu),            /* 135 - .. - .. */
                                            ((cr2 * 13u) >> 4) & 0x00ff00ffu);
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const unsigned w = yw[r][k];
                    const unsigned te = ((mvg_pair_even(w) * 149u) >> 7) & 0x01ff01ffu;       /* pixels 4k, 4k+2 */
                    const unsigned to = ((mvg_pair_odd(w) * 149u) >> 7) & 0x01ff01ffu;        /* pixels 4k+1, 4k+3 */
                    /* chroma sample 2k serves pixels 4k, 4k+1; sample 2k+1 serves 4k+2, 4k+3: both pairs line up */
                    const unsigned Re = mvg_add_clip8x2(te, rC), Ro = mvg_add_clip8x2(to, rC);
                    const unsigned Ge = mvg_add_clip8x2(te, gC), Go = mvg_add_clip8x2(to, gC);
                    const unsigned Be = mvg_add_clip8x2(te, bC), Bo = mvg_add_clip8x2(to, bC);
                    const unsigned X = __byte_perm(Re, Ge, 0x6240);       /* R0 G0 R2 G2 */
                    const unsigned Y = __byte_perm(Be, Ro, 0x6240);       /* B0 R1 B2 R3 */
                    const unsigned Z = __byte_perm(Go, Bo, 0x6240);       /* G1 B1 G3 B3 */
                    out[r][3 * k]     = __byte_perm(X, Y, 0x5410);        /* R0 G0 B0 R1 */
                    out[r][3 * k + 1] = __byte_perm(Z, X, 0x7610);        /* G1 B1 R2 G2 */
                    out[r][3 * k + 2] = __byte_perm(Y, Z, 0x7632);        /* B2 R3 G3 B3 */
                }
            }
#pragma unroll
            for (int r = 0; r < 2; r++) {
                uint4 *dst = reinterpret_cast<uint4 *>(rgb + (size_t)pic * ysz * 3 +
                                                       ((size_t)(my * 16 + 2 * q + r) * p.width + mx * 16) * 3);
                dst[0] = make_uint4(out[r][0], out[r][1], out[r][2], out[r][3]);
                dst[1] = make_uint4(out[r][4], out[r][5], out[r][6], out[r][7]);
                dst[2] = make_uint4(out[r][8], out[r][9], out[r][10], out[r][11]);
            }
        }
        mx += dmx; my += dmy; pic += dpic;
        if (mx >= w_mbs) { mx -= w_mbs; my++; }
        if (my >= h_mbs) { my -= h_mbs; pic++; }
        __syncthreads();
    }
}

/* scale s > 1, any s that divides the picture: one thread per output pixel, rounded s x s box average of the
 * full-resolution RGB picture.  Slow (byte loads); k3_rgb_scaled covers the
 * power-of-two scales up to 16. */
__global__ void __launch_bounds__(256)
k3_rgb_scaled_generic(K3Params p)
{
    const int s = p.scale, ow = p.width / s, oh = p.height / s;
    const long long per_pic = (long long)ow * oh, total = per_pic * p.n_pics;
    const size_t ysz = (size_t)p.width * p.height;
    const int area = s * s, w_mbs = p.width >> 4;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int pic = (int)(g / per_pic);
        const long long rem = g - (long long)pic * per_pic;
        const int oy = (int)(rem / ow), ox = (int)(rem - (long long)oy * ow);
        const size_t slot = (size_t)(p.first_slot + pic);
        const uint8_t *tiles = p.tiles + slot * (ysz * 3 / 2);
        int aR = 0, aG = 0, aB = 0;
        for (int dy = 0; dy < s; dy++) {
            const int py = oy * s + dy;
            for (int dx = 0; dx < s; dx++) {
                const int px = ox * s + dx;
                const uint8_t *tile = tiles + ((size_t)(py >> 4) * w_mbs + (px >> 4)) * 384;
                const int co = ((py & 15) >> 1) * 8 + ((px & 15) >> 1);
                int R, G, B;
                mvg_ycc_to_rgb(__ldg(tile + (py & 15) * 16 + (px & 15)), __ldg(tile + 256 + co), __ldg(tile + 320 + co), R, G, B);
                aR += R; aG += G; aB += B;
            }
        }
        uint8_t *o = p.rgb + slot * ((size_t)ow * oh * 3) + ((size_t)oy * ow + ox) * 3;
        o[0] = (uint8_t)((aR + area / 2) / area);
        o[1] = (uint8_t)((aG + area / 2) / area);
        o[2] = (uint8_t)((aB + area / 2) / area);
    }
}

/* scale s in {2, 4, 8, 16}: rounded s x s box average of the scale-1 RGB picture.
 * Same staging as k3_rgb_full (32 tiles per CTA in shared memory); a thread then converts one 4x4 cell of a
 * macroblock -- four 2x2 quads, each with its own chroma sample -- with the int16-pair arithmetic of
 * k3_rgb_full and keeps per-quad colour sums.  s = 2: a quad is an output pixel; s = 4: the cell is; s = 8, 16:
 * the cells of an output pixel sit in adjacent lanes (cells are numbered in Z order) and are added with shuffles. */
__global__ void __launch_bounds__(256)
k3_rgb_scaled(K3Params p)
{
    __shared__ __align__(16) uint8_t s_tiles[K3_MBS * K3_TILE];
    const int w_mbs = p.width >> 4, h_mbs = p.height >> 4, n_mb = w_mbs * h_mbs;
    const long long total_mbs = (long long)n_mb * p.n_pics;
    const long long n_groups = (total_mbs + K3_MBS - 1) / K3_MBS;
    const int sc = p.scale, ow = p.width / sc, oh = p.height / sc;
    const int sh = sc == 2 ? 2 : sc == 4 ? 4 : sc == 8 ? 6 : 8;        /* log2(s * s) */
    const size_t out_pic = (size_t)ow * oh * 3;
    const uint8_t *tiles = p.tiles + (size_t)p.first_slot * n_mb * 384;
    uint8_t *rgb = p.rgb + (size_t)p.first_slot * out_pic;
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long long first = g * K3_MBS;
        const int n_here = (int)min((long long)K3_MBS, total_mbs - first);
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(tiles + (size_t)first * 384);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int ch = threadIdx.x + 256 * k;
                if (ch < n_here * 24) {
                    const int t = ch / 24;
                    *reinterpret_cast<uint4 *>(s_tiles + t * K3_TILE + (ch - t * 24) * 16) = __ldg(src + ch);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int cell = threadIdx.x + 256 * half;              /* 16 cells per macroblock, Z order */
            const int mbl = cell >> 4, cz = cell & 15;
            const int cx = (cz & 1) | ((cz >> 1) & 2), cy = ((cz >> 1) & 1) | ((cz >> 2) & 2);
            const bool live = mbl < n_here;
            unsigned sumR = 0, sumG = 0, sumB = 0;                  /* s >= 4: int16 pairs {even columns, odd columns} */
            unsigned q2[2][2][3];                                   /* s == 2: per quad */
            if (live) {
                const uint8_t *tile = s_tiles + mbl * K3_TILE;
                const unsigned cbp = *reinterpret_cast<const uint16_t *>(tile + 256 + (cy * 2) * 8 + cx * 2) |
                                     (unsigned)*reinterpret_cast<const uint16_t *>(tile + 256 + (cy * 2 + 1) * 8 + cx * 2) << 16;
                const unsigned crp = *reinterpret_cast<const uint16_t *>(tile + 320 + (cy * 2) * 8 + cx * 2) |
                                     (unsigned)*reinterpret_cast<const uint16_t *>(tile + 320 + (cy * 2 + 1) * 8 + cx * 2) << 16;
#pragma unroll
                for (int qy = 0; qy < 2; qy++) {
                    /* chroma samples (2cx, 2cy+qy) and (2cx+1, 2cy+qy) as an int16 pair: pixels x and x+2 of a row */
                    const unsigned cb2 = qy ? mvg_pair_hi(cbp) : mvg_pair_lo(cbp);
                    const unsigned cr2 = qy ? mvg_pair_hi(crp) : mvg_pair_lo(crp);
                    const unsigned rC = __vsub2(((cr2 * 204u) >> 7) & 0x01ff01ffu, 0x00de00deu);
                    const unsigned bC = __vsub2(((cb2 * 129u) >> 6) & 0x03ff03ffu, 0x01140114u);
                    const unsigned gC = __vsub2(__vsub2(0x00870087u, ((cb2 * 25u) >> 6) & 0x00ff00ffu), ((cr2 * 13u) >> 4) & 0x00ff00ffu);
                    unsigned aR = 0, aG = 0, aB = 0;                /* halves: quad qx = 0 (pixels 0,1) | quad qx = 1 (pixels 2,3) */
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        const unsigned w = *reinterpret_cast<const unsigned *>(tile + (cy * 4 + qy * 2 + r) * 16 + cx * 4);
                        const unsigned te = ((mvg_pair_even(w) * 149u) >> 7) & 0x01ff01ffu;       /* pixels 0, 2 */
                        const unsigned to = ((mvg_pair_odd(w) * 149u) >> 7) & 0x01ff01ffu;        /* pixels 1, 3 */
                        aR += mvg_add_clip8x2(te, rC) + mvg_add_clip8x2(to, rC);
                        aG += mvg_add_clip8x2(te, gC) + mvg_add_clip8x2(to, gC);
                        aB += mvg_add_clip8x2(te, bC) + mvg_add_clip8x2(to, bC);
                    }
                    q2[qy][0][0] = aR & 0xffffu; q2[qy][1][0] = aR >> 16;
                    q2[qy][0][1] = aG & 0xffffu; q2[qy][1][1] = aG >> 16;
                    q2[qy][0][2] = aB & 0xffffu; q2[qy][1][2] = aB >> 16;
                    sumR += aR; sumG += aG; sumB += aB;
                }
            }
            /* position of the macroblock in the picture */
            const long long mb = first + mbl;
            const int pic = (int)(mb / n_mb), rem = (int)(mb - (long long)pic * n_mb), my = rem / w_mbs, mx = rem - my * w_mbs;
            uint8_t *out = rgb + (size_t)pic * out_pic;
            if (sc == 2) {
                if (live) {
#pragma unroll
                    for (int qy = 0; qy < 2; qy++)
#pragma unroll
                        for (int qx = 0; qx < 2; qx++) {
                            uint8_t *o = out + ((size_t)(my * 8 + cy * 2 + qy) * ow + mx * 8 + cx * 2 + qx) * 3;
                            o[0] = (uint8_t)((q2[qy][qx][0] + 2) >> 2);
                            o[1] = (uint8_t)((q2[qy][qx][1] + 2) >> 2);
                            o[2] = (uint8_t)((q2[qy][qx][2] + 2) >> 2);
                        }
                }
            } else {
                unsigned R = (sumR & 0xffffu) + (sumR >> 16), G = (sumG & 0xffffu) + (sumG >> 16), B = (sumB & 0xffffu) + (sumB >> 16);
                const int span = sc == 4 ? 1 : sc == 8 ? 4 : 16;    /* cells per output pixel: adjacent lanes */
                for (int o = 1; o < span; o <<= 1) {
                    R += __shfl_xor_sync(MVG_FULL, R, o); G += __shfl_xor_sync(MVG_FULL, G, o); B += __shfl_xor_sync(MVG_FULL, B, o);
                }
                if (live && (cz & (span - 1)) == 0) {
                    const int per = 16 / sc;                        /* output pixels per macroblock side */
                    const int ox = sc == 4 ? cx : sc == 8 ? cx >> 1 : 0, oy = sc == 4 ? cy : sc == 8 ? cy >> 1 : 0;
                    uint8_t *o = out + ((size_t)(my * per + oy) * ow + mx * per + ox) * 3;
                    const unsigned rnd = 1u << (sh - 1);
                    o[0] = (uint8_t)((R + rnd) >> sh);
                    o[1] = (uint8_t)((G + rnd) >> sh);
                    o[2] = (uint8_t)((B + rnd) >> sh);
                }
            }
        }
        __syncthreads();
    }
}

/* ========================================================================= */
/* Kernel 4: tiles -> planar I420, the gather of export_idr_yuv420() (export.c:100-146).  Only runs when
 * a caller asks for planar YUV; the RGB path reads the tiles directly, as mb_to_rgb() does.
 * One thread moves two luma rows of a macroblock (32 contiguous bytes in) or, for the last quarter of
 * the index space, two rows of each chroma plane; neighbouring threads take neighbouring macroblocks. */
__global__ void __launch_bounds__(256)
k4_yuv_planar(K3Params p)
{
    const int w_mbs = p.width >> 4, h_mbs = p.height >> 4;
    const long long per_pic = (long long)w_mbs * h_mbs * 12;       /* 8 luma row pairs + 4 chroma row pairs per MB */
    const long long total = per_pic * p.n_pics;
    const size_t ysz = (size_t)p.width * p.height;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int pic = (int)(g / per_pic);
        const int rem = (int)(g - (long long)pic * per_pic);
        const int line = rem / w_mbs, mx = rem - line * w_mbs;     /* line: my * 12 + piece */
        const int my = line / 12, piece = line - my * 12;
        const size_t slot = (size_t)(p.first_slot + pic);
        const uint8_t *tile = p.tiles + slot * (ysz * 3 / 2) + ((size_t)my * w_mbs + mx) * 384;
        uint8_t *Y = p.yuv + slot * (ysz * 3 / 2);
        if (piece < 8) {
            const uint4 a = __ldg(reinterpret_cast<const uint4 *>(tile + piece * 32));
            const uint4 b = __ldg(reinterpret_cast<const uint4 *>(tile + piece * 32 + 16));
            uint8_t *d = Y + (size_t)(my * 16 + piece * 2) * p.width + mx * 16;
            *reinterpret_cast<uint4 *>(d) = a;
            *reinterpret_cast<uint4 *>(d + p.width) = b;
        } else {
            const int q = piece - 8;                               /* chroma rows 2q, 2q+1 of both planes */
            const uint4 cb = __ldg(reinterpret_cast<const uint4 *>(tile + 256 + q * 16));
            const uint4 cr = __ldg(reinterpret_cast<const uint4 *>(tile + 320 + q * 16));
            const int cw = p.width >> 1;
            uint8_t *d = Y + ysz + (size_t)(my * 8 + q * 2) * cw + mx * 8;
            *reinterpret_cast<uint2 *>(d) = make_uint2(cb.x, cb.y);
            *reinterpret_cast<uint2 *>(d + cw) = make_uint2(cb.z, cb.w);
            d += ysz / 4;
            *reinterpret_cast<uint2 *>(d) = make_uint2(cr.x, cr.y);
            *reinterpret_cast<uint2 *>(d + cw) = make_uint2(cr.z, cr.w);
        }
    }
}
