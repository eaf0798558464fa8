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
#include <stdint.h>

#include "mvg_internal.h"

#define MVG_FULL 0xffffffffu

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

/* Table 8-15: QPC as a function of qPI >= 30 (h264_transform.c:71) */
__constant__ unsigned char mvg_qpc_tab[22] = {29,30,31,32,32,33,34,34,35,35,36,36,37,37,37,38,38,38,39,39,39,39};

/* h264_transform.c:598-637 (8-bit video: QpBdOffsetC = 0) */
__device__ __forceinline__ int mvg_chroma_qp(int qp_y, int offset)
{
    const int qpi = min(max(qp_y + offset, 0), 51);
    return qpi < 30 ? qpi : (int)mvg_qpc_tab[qpi - 30];
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

#define K1_WARPS 4          /* warps per CTA                      */
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

struct K1WarpSmem {
    __align__(128) int16_t tile[2][K1_TILE];    /* levels in, residual out (in place), double buffered */
    int32_t  dc[K1_GROUP][24];                  /* dequantised DC of Intra16x16 luma / chroma blocks   */
    int32_t  f1[K1_GROUP][16];                  /* first stage of the luma DC Hadamard                 */
    int32_t  tr[4][8][9];                       /* 8x8 transpose, padded                               */
    uint32_t meta[K1_GROUP];                    /* mb_kind | QPY << 8                                  */
    uint8_t  list4[K1_GROUP * 24 + 8];          /* 4x4 blocks that need the full transform             */
    uint8_t  list8[K1_GROUP * 4 + 4];           /* 8x8 blocks with non-zero levels                     */
    __align__(8) uint64_t mbar[2];
};

/* position (bx,by) -> luma4x4BlkIdx (h264_spatial.c:210-225 inverted) */
__device__ __forceinline__ int mvg_blk_of(int bx, int by) { return (bx & 1) | ((by & 1) << 1) | ((bx >> 1) << 2) | ((by >> 1) << 3); }

/* One warp transforms K1_GROUP macroblocks per iteration, in place in shared memory:
 *   - the 3 KB of levels arrive with ONE bulk asynchronous copy (cp.async.bulk + mbarrier), the
 *     next group's copy is in flight while this one is processed, and the residual leaves with
 *     one bulk store -- no per-thread global loads/stores for the payload;
 *   - Intra16x16 luma DC (4x4 Hadamard, h264_transform.c:756-812) and chroma DC (2x2, :827-936)
 *     are done first, separably, a few lanes per macroblock;
 *   - every 4x4 block is classified: all-zero (nothing to do), DC-only (every residual sample is
 *     (d00 + 32) >> 6, spec 8.5.12.2 with a single non-zero input) or general; the general blocks
 *     of the whole group are compacted with ballots and run through quant4x4 + idct4x4
 *     (h264_transform.c:1100-1191) 32 at a time, one block per lane, entirely in registers;
 *   - non-zero 8x8 blocks (Intra8x8 luma, quant8x8/idct8x8 :1256-1383) are compacted the same
 *     way and transformed 4 per pass, 8 lanes per block, transposed through shared memory.
 * Residual layout out: per macroblock 24 blocks x 16 int16, block-major (see MvgMbCtl). */
__global__ void __launch_bounds__(K1_WARPS * 32, 5)
k1_dequant_idct(K1Params p)
{
    __shared__ int32_t s_ls4[3 * 6 * 16];
    __shared__ int32_t s_ls4q[3 * 52 * 16];                     /* per qP; << (qP/6-4) folded in when qP >= 24 */
    __shared__ int32_t s_ls8[6 * 64];
    __shared__ uint8_t s_zz8inv[64];
    __shared__ K1WarpSmem s_warp[K1_WARPS];

    for (int i = threadIdx.x; i < 3 * 6 * 16; i += blockDim.x) s_ls4[i] = (&p.tab->ls4[0][0][0])[i];
    for (int i = threadIdx.x; i < 3 * 52 * 16; i += blockDim.x) s_ls4q[i] = (&p.tab->ls4q[0][0][0])[i];
    for (int i = threadIdx.x; i < 6 * 64; i += blockDim.x) s_ls8[i] = (&p.tab->ls8[0][0])[i];
    if (threadIdx.x < 64) s_zz8inv[threadIdx.x] = p.tab->zz8inv[threadIdx.x];
    const int cb_off = p.tab->cb_qp_offset, cr_off = p.tab->cr_qp_offset;

    const int lane = threadIdx.x & 31;
    K1WarpSmem &s = s_warp[threadIdx.x >> 5];
    if (lane == 0) { mvg_mbar_init(&s.mbar[0], 1); mvg_mbar_init(&s.mbar[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const long long n_groups = (p.n_mbs + K1_GROUP - 1) / K1_GROUP;
    const long long stride = (long long)gridDim.x * K1_WARPS;
    long long g = (long long)blockIdx.x * K1_WARPS + (threadIdx.x >> 5);

    /* side information of a group, one 32-bit word per lane: lane = 8*j + t for macroblock j;
     * t = 0..3 luma modes 4t..4t+3, t = 4 mb_kind, 5 QPY, 6 Intra16x16PredMode, 7 intra_chroma_pred_mode */
    const int mj = lane >> 3, mt = lane & 7;
    auto load_meta = [&](long long grp) -> unsigned {
        const long long mb = grp * K1_GROUP + mj;
        if (mb >= p.n_mbs) return 0u;
        if (mt < 4) return __ldg(reinterpret_cast<const unsigned *>(p.luma_modes + mb * 16) + mt);
        const uint8_t *src = mt == 4 ? p.mb_kind : mt == 5 ? reinterpret_cast<const uint8_t *>(p.qp_y)
                           : mt == 6 ? p.i16_mode : p.chroma_mode;
        return (unsigned)__ldg(src + mb);
    };
    auto issue_load = [&](int buf, long long grp) {
        const long long mb0 = grp * K1_GROUP;
        const unsigned bytes = (unsigned)min((long long)K1_GROUP, p.n_mbs - mb0) * 768u;
        mvg_mbar_expect_tx(&s.mbar[buf], bytes);
        mvg_bulk_load(s.tile[buf], p.coeff + mb0 * 384, bytes, &s.mbar[buf]);
    };

    unsigned nmeta = 0;
    if (g < n_groups) {
        if (lane == 0) issue_load(0, g);
        nmeta = load_meta(g);
    }
    unsigned parity = 0;            /* bit b: phase parity of mbar[b] */

    for (int it = 0; g < n_groups; g += stride, it++) {
        const int buf = it & 1;
        const unsigned meta = nmeta;
        {   /* next group: its buffer was the source of the previous bulk store */
            const long long gn = g + stride;
            if (gn < n_groups) {
                if (lane == 0) { mvg_bulk_wait_read(); issue_load(buf ^ 1, gn); }
                nmeta = load_meta(gn);
            }
        }
        const long long mb0 = g * K1_GROUP;
        const int nmb = (int)min((long long)K1_GROUP, p.n_mbs - mb0);
        int16_t *tile = s.tile[buf];

        /* per-lane view of "my" macroblock j = lane >> 3 */
        const int kind_j = (int)__shfl_sync(MVG_FULL, meta, (lane & ~7) + 4);
        const int qp_j = (signed char)__shfl_sync(MVG_FULL, meta, (lane & ~7) + 5);
        if (mt == 0) s.meta[mj] = (unsigned)kind_j | ((unsigned)(qp_j & 255) << 8);

        mvg_mbar_wait(&s.mbar[buf], (parity >> buf) & 1u);
        parity ^= 1u << buf;

        /* ---------------- DC transforms ---------------- */
        if (mj < nmb) {
            const int16_t *cf = tile + mj * 384;
            if (kind_j == MVG_MB_I16x16 && mt < 4) {         /* row mt of c: t = c * H */
                const int a = cf[mvg_blk_of(0, mt) * 16], b = cf[mvg_blk_of(1, mt) * 16];
                const int c = cf[mvg_blk_of(2, mt) * 16], d = cf[mvg_blk_of(3, mt) * 16];
                int32_t *f1 = s.f1[mj] + mt * 4;
                f1[0] = a + b + c + d; f1[1] = a + b - c - d; f1[2] = a - b - c + d; f1[3] = a - b + c - d;
            } else if (mt == 4 || mt == 5) {                 /* chroma plane mt-4: f = A c A, then scale */
                const int pl = mt - 4;
                const int16_t *cc = cf + 256 + pl * 64;
                const int c00 = cc[0], c01 = cc[16], c10 = cc[32], c11 = cc[48];
                const int qpc = mvg_chroma_qp(qp_j, pl ? cr_off : cb_off);
                const int qd = qpc / 6, ls00 = s_ls4[((pl + 1) * 6 + (qpc - 6 * qd)) * 16];
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
            const int qd = qp_j / 6, ls00 = s_ls4[(qp_j - 6 * qd) * 16];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int t = f[i] * ls00;
                s.dc[mj][mvg_blk_of(mt, i)] = (qp_j >= 36) ? (int)((unsigned)t << (qd - 6)) : ((t + (1 << (5 - qd))) >> (6 - qd));
            }
        }
        __syncwarp();

        /* ---------------- classify the 4x4 blocks, compact the general ones ---------------- */
        int n4 = 0, n8 = 0;
        unsigned nzbal[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int u = lane + 32 * r, j = u / 24, b = u - 24 * j;
            bool general = false, nonzero = false, nz8q = false;
            const unsigned mw = j < nmb ? s.meta[j] : 0u;
            const int kind = mw & 255, qp = (signed char)(mw >> 8);
            const bool is8 = kind == MVG_MB_I8x8 && b < 16;
            if (j < nmb) {
                uint4 *blk = reinterpret_cast<uint4 *>(tile + j * 384 + b * 16);
                const uint4 w0 = blk[0], w1 = blk[1];
                const unsigned rest = (w0.x & 0xffff0000u) | w0.y | w0.z | w0.w | w1.x | w1.y | w1.z | w1.w;
                const int dcraw = (short)(w0.x & 0xffff);
                if (is8) nz8q = (rest | (unsigned)dcraw) != 0;
                else if (rest) general = nonzero = true;
                else {
                    const bool has_dc = b >= 16 || kind == MVG_MB_I16x16;
                    int rv = 0;
                    if (has_dc) rv = (s.dc[j][b] + 32) >> 6;          /* d00 = c00 (h264_transform.c:1126-1129) */
                    else if (dcraw) {
                        const int qd = qp / 6, lq = s_ls4q[qp * 16];
                        const int d = qp > 23 ? dcraw * lq : (dcraw * lq + (1 << (3 - qd))) >> (4 - qd);
                        rv = (d + 32) >> 6;
                    }
                    rv = min(max(rv, -32768), 32767);
                    nonzero = rv != 0;
                    if (nonzero || dcraw) {
                        const unsigned pk = (unsigned)(rv & 0xffff) * 0x10001u;
                        blk[0] = make_uint4(pk, pk, pk, pk); blk[1] = make_uint4(pk, pk, pk, pk);
                    }
                }
            }
            const unsigned gb = __ballot_sync(MVG_FULL, general);
            if (general) s.list4[n4 + __popc(gb & ((1u << lane) - 1))] = (uint8_t)u;
            n4 += __popc(gb);
            /* Intra8x8: slots 4*b8..4*b8+3 are the four quarters of 8x8 block b8 (aligned lane quads) */
            const unsigned qb = __ballot_sync(MVG_FULL, nz8q);
            const bool any8 = is8 && ((qb >> (lane & ~3)) & 0xFu) != 0;
            const bool lead8 = any8 && (lane & 3) == 0;
            const unsigned lb = __ballot_sync(MVG_FULL, lead8);
            if (lead8) s.list8[n8 + __popc(lb & ((1u << lane) - 1))] = (uint8_t)(j * 4 + (b >> 2));
            n8 += __popc(lb);
            nzbal[r] = __ballot_sync(MVG_FULL, nonzero || any8);
        }
        __syncwarp();

        /* ---------------- general 4x4 blocks, 32 per pass ---------------- */
        for (int base = 0; base < n4; base += 32) {
            if (base + lane < n4) {
                const int u = s.list4[base + lane], j = u / 24, b = u - 24 * j;
                const unsigned mw = s.meta[j];
                const int kind = mw & 255, qp = (signed char)(mw >> 8);
                const int comp = b < 16 ? 0 : (b < 20 ? 1 : 2);
                const int qpb = comp ? mvg_chroma_qp(qp, comp == 1 ? cb_off : cr_off) : qp;
                uint4 *blk = reinterpret_cast<uint4 *>(tile + j * 384 + b * 16);
                const uint4 a = blk[0], bb = blk[1];
                int c[16];                      /* zig-zag k -> (row,col): utils.h:64 / spec Table 8-13 */
                c[0] = (short)(a.x & 0xffff); c[1] = (int)a.x >> 16; c[4] = (short)(a.y & 0xffff); c[8] = (int)a.y >> 16;
                c[5] = (short)(a.z & 0xffff); c[2] = (int)a.z >> 16; c[3] = (short)(a.w & 0xffff); c[6] = (int)a.w >> 16;
                c[9] = (short)(bb.x & 0xffff); c[12] = (int)bb.x >> 16; c[13] = (short)(bb.y & 0xffff); c[10] = (int)bb.y >> 16;
                c[7] = (short)(bb.z & 0xffff); c[11] = (int)bb.z >> 16; c[14] = (short)(bb.w & 0xffff); c[15] = (int)bb.w >> 16;
                const bool keep_dc = comp != 0 || kind == MVG_MB_I16x16;
                const int dc_in = keep_dc ? s.dc[j][b] : 0;
                /* quant4x4, h264_transform.c:1100-1134 */
                const int4 *lq = reinterpret_cast<const int4 *>(s_ls4q + (comp * 52 + qpb) * 16);
                const int4 l0 = lq[0], l1 = lq[1], l2 = lq[2], l3 = lq[3];
                const int ls[16] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w, l2.x, l2.y, l2.z, l2.w, l3.x, l3.y, l3.z, l3.w};
                if (qpb > 23) {
#pragma unroll
                    for (int k = 0; k < 16; k++) c[k] = c[k] * ls[k];
                } else {
                    const int qd = qpb / 6, rnd = 1 << (3 - qd), sh = 4 - qd;
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
                o0.x = mvg_pack_sat16(c[1] >> 6, c[0] >> 6);   o0.y = mvg_pack_sat16(c[3] >> 6, c[2] >> 6);
                o0.z = mvg_pack_sat16(c[5] >> 6, c[4] >> 6);   o0.w = mvg_pack_sat16(c[7] >> 6, c[6] >> 6);
                o1.x = mvg_pack_sat16(c[9] >> 6, c[8] >> 6);   o1.y = mvg_pack_sat16(c[11] >> 6, c[10] >> 6);
                o1.z = mvg_pack_sat16(c[13] >> 6, c[12] >> 6); o1.w = mvg_pack_sat16(c[15] >> 6, c[14] >> 6);
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
                const int qp = (signed char)(s.meta[j] >> 8);
                const int qd8 = qp / 6;
                const int32_t *l8 = s_ls8 + (qp - 6 * qd8) * 64 + row * 8;
                const int16_t *in = tile + j * 384 + b8 * 64;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int t = (int)in[s_zz8inv[row * 8 + q]] * l8[q];       /* quant8x8, h264_transform.c:1256-1284 */
                    v[q] = (qp > 35) ? (int)((unsigned)t << (qd8 - 6)) : ((t + (1 << (5 - qd8))) >> (6 - qd8));
                }
                if (row == 0) v[0] += 32;                                       /* rounding of the final >> 6 (:1382) */
                mvg_idct8_1d(v);                                                /* row pass */
#pragma unroll
                for (int q = 0; q < 8; q++) s.tr[lane >> 3][row][q] = v[q];
                o8 = tile + j * 384 + (b8 * 4 + (row >> 2)) * 16 + (row & 3);
            }
            __syncwarp();
            if (act) {
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = s.tr[lane >> 3][i][row];     /* this lane now owns column `row` */
                mvg_idct8_1d(v);                                                /* column pass */
#pragma unroll
                for (int i = 0; i < 8; i++) o8[(i >> 2) * 32 + (i & 3) * 4] = (int16_t)min(max(v[i] >> 6, -32768), 32767);
            }
            __syncwarp();
        }
        __syncwarp();

        /* ---------------- residual out (one bulk store) + control records ---------------- */
        if (lane == 0) mvg_bulk_store(p.resid + mb0 * 384, tile, (unsigned)nmb * 768u);
        {
            /* 96-bit non-zero map -> 24 bits per macroblock */
            const unsigned long long lo = (unsigned long long)nzbal[0] | ((unsigned long long)nzbal[1] << 32);
            const unsigned hi = nzbal[2];
            const unsigned nz = mj == 0 ? nzbal[0] : mj == 1 ? (unsigned)(lo >> 24)
                              : mj == 2 ? ((unsigned)(lo >> 48) | (hi << 16)) : (hi >> 8);
            /* 4 mode bytes -> 4 nibbles; lane 8j collects its macroblock's record */
            unsigned nib = (meta & 0xF) | ((meta >> 4) & 0xF0) | ((meta >> 8) & 0xF00) | ((meta >> 12) & 0xF000);
            const unsigned n1 = __shfl_down_sync(MVG_FULL, nib, 1), n2 = __shfl_down_sync(MVG_FULL, nib, 2);
            const unsigned n3 = __shfl_down_sync(MVG_FULL, nib, 3);
            const unsigned k4 = __shfl_down_sync(MVG_FULL, meta, 4), k6 = __shfl_down_sync(MVG_FULL, meta, 6);
            const unsigned k7 = __shfl_down_sync(MVG_FULL, meta, 7);
            if (mt == 0 && mj < nmb)
                *reinterpret_cast<uint4 *>(p.ctl + mb0 + mj) =
                    make_uint4((k4 & 255) | ((k6 & 255) << 8) | ((k7 & 255) << 16), nib | (n1 << 16), n2 | (n3 << 16), nz & 0x00FFFFFFu);
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
    uint8_t        *yuv;        /* [slot][1.5*W*H] planar I420                */
    uint2          *halo;       /* [slot][h_mbs][w_mbs][8]: bottom sample row of every MB,
                                   4 data bytes + 4 flag bytes per 64-bit word  */
    int            *work;       /* work counter of this launch (starts at 0)  */
    const MvgLuts  *luts;
    unsigned        epoch;      /* flag value that marks halo words of THIS launch */
    int w_mbs, h_mbs, first_slot, n_pics, group;
};

#define K2_WARPS 4

struct K2WarpSmem {
    __align__(16) int16_t  resid[384];
    __align__(16) uint8_t  lt[MVG_LT_ROWS * MVG_LT_STRIDE];
    __align__(16) uint8_t  ct[2][MVG_CT_ROWS * MVG_CT_STRIDE];
    __align__(16) uint32_t n8[32];      /* Intra8x8 neighbour line: p' | f2 << 8 | f3 << 16 per entry */
};

__device__ __forceinline__ uint2 mvg_ld_relaxed_u64(const uint2 *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return make_uint2((unsigned)v, (unsigned)(v >> 32));
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

/* ---- Intra16x16 luma (h264_intra_prediction.c:1945-2141) ------------------- */
__device__ __forceinline__ void k2_luma16(K2WarpSmem &s, int lane, int mode, bool left, bool up)
{
    uint8_t *lt = s.lt;
    const int y = lane >> 1, x0 = (lane & 1) * 8;
    int pred[8];
    if (mode == 0) {            /* Vertical */
        const uint2 t = *reinterpret_cast<const uint2 *>(lt + MVG_LT_XOFF + x0);
#pragma unroll
        for (int i = 0; i < 4; i++) { pred[i] = (t.x >> (8 * i)) & 255; pred[4 + i] = (t.y >> (8 * i)) & 255; }
    } else if (mode == 1) {     /* Horizontal */
        const int v = lt[(y + 1) * MVG_LT_STRIDE + MVG_LT_XOFF - 1];
#pragma unroll
        for (int i = 0; i < 8; i++) pred[i] = v;
    } else if (mode == 2) {     /* DC */
        int v = 0;
        if (lane < 16) {
            if (up) v += lt[MVG_LT_XOFF + lane];
            if (left) v += lt[(lane + 1) * MVG_LT_STRIDE + MVG_LT_XOFF - 1];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MVG_FULL, v, o);
        v = (left && up) ? (v + 16) >> 5 : (left || up) ? (v + 8) >> 4 : 128;
#pragma unroll
        for (int i = 0; i < 8; i++) pred[i] = v;
    } else {                    /* Plane */
        int term = 0;
        const int i = lane & 7;
        if (lane < 8)       term = (i + 1) * ((int)lt[MVG_LT_XOFF + 8 + i] - (int)lt[MVG_LT_XOFF + 6 - i]);
        else if (lane < 16) term = (i + 1) * ((int)lt[(9 + i) * MVG_LT_STRIDE + MVG_LT_XOFF - 1] -
                                              (int)lt[(7 - i) * MVG_LT_STRIDE + MVG_LT_XOFF - 1]);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) term += __shfl_xor_sync(MVG_FULL, term, o);
        const int H = __shfl_sync(MVG_FULL, term, 0), V = __shfl_sync(MVG_FULL, term, 8);
        const int a = 16 * ((int)lt[16 * MVG_LT_STRIDE + MVG_LT_XOFF - 1] + (int)lt[MVG_LT_XOFF + 15]);
        const int b = (5 * H + 32) >> 6, c = (5 * V + 32) >> 6;
        const int base = a + c * (y - 7) + 16;
#pragma unroll
        for (int k = 0; k < 8; k++) pred[k] = mvg_clip8((base + b * (x0 + k - 7)) >> 5);
    }
    /* row y, samples x0..x0+7: 4x4 blocks (x0/4, y/4) and (x0/4+1, y/4), row y&3 of each */
    const int bcol = x0 >> 2, brow = y >> 2;
    const int blkA = (bcol & 1) | ((brow & 1) << 1) | ((bcol >> 1) << 2) | ((brow >> 1) << 3);
    const uint2 ra = *reinterpret_cast<const uint2 *>(s.resid + blkA * 16 + (y & 3) * 4);
    const uint2 rb = *reinterpret_cast<const uint2 *>(s.resid + (blkA + 1) * 16 + (y & 3) * 4);
    const unsigned rr[4] = {ra.x, ra.y, rb.x, rb.y};
    unsigned lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int r0 = (short)(rr[k] & 0xffff), r1 = (int)rr[k] >> 16;
        const unsigned p0 = (unsigned)mvg_add_clip8(pred[2 * k], r0), p1 = (unsigned)mvg_add_clip8(pred[2 * k + 1], r1);
        const unsigned pk = p0 | (p1 << 8);
        if (k < 2) lo |= pk << (16 * k); else hi |= pk << (16 * (k - 2));
    }
    *reinterpret_cast<uint2 *>(lt + (y + 1) * MVG_LT_STRIDE + MVG_LT_XOFF + x0) = make_uint2(lo, hi);
}

/* ---- Intra4x4 luma: anti-diagonal schedule, two blocks per step ------------- */
/* Blocks with bx + 2*by == t can be predicted together: their left, up, up-left and
 * up-right neighbours all belong to earlier steps.  Lanes 0..15 take block
 * (t&1, t>>1), lanes 16..31 block ((t&1)+2, (t>>1)-1), one sample per lane. */
template <int T>
__device__ __forceinline__ void k2_luma4_step(K2WarpSmem &s, const uint32_t *lut4, int half, int pix, int pxy,
                                              unsigned mlo, unsigned mhi, bool availA, bool availB, bool availC)
{
    constexpr int bx0 = (T & 1), by0 = (T >> 1), bx1 = (T & 1) + 2, by1 = (T >> 1) - 1;
    constexpr bool v0 = by0 <= 3, v1 = by1 >= 0 && by1 <= 3;
    constexpr int blk0 = (bx0 & 1) | ((by0 & 1) << 1) | ((bx0 >> 1) << 2) | ((by0 >> 1) << 3);
    constexpr int blk1 = (bx1 & 1) | ((by1 & 1) << 1) | ((bx1 >> 1) << 2) | (((by1 >> 1) & 1) << 3);
    uint8_t *lt = s.lt;
    const bool valid = half ? v1 : v0;
    if (valid) {
        const int bx = half ? bx1 : bx0, by = half ? by1 : by0, blk = half ? blk1 : blk0;
        const unsigned mw = blk < 8 ? mlo : mhi;
        const int mode = (mw >> (4 * (blk & 7))) & 15;
        const bool left = bx > 0 || availA, up = by > 0 || availB;
        /* p[4..7,-1]: h264_intra_prediction.c:398-429 + h264_spatial.c:757-774 */
        bool tr;
        if (blk == 3 || blk == 11) tr = false;
        else if (by > 0) tr = bx < 3;
        else tr = bx < 3 ? availB : availC;
        const int org = (by * 4 + 1) * MVG_LT_STRIDE + MVG_LT_XOFF + bx * 4;
        int pred;
        if (mode == 2) {
            int sum = 0;
            if (up) sum += mvg_sum4(*reinterpret_cast<const unsigned *>(lt + org - MVG_LT_STRIDE));
            if (left) sum += (int)lt[org - 1] + (int)lt[org + MVG_LT_STRIDE - 1] +
                             (int)lt[org + 2 * MVG_LT_STRIDE - 1] + (int)lt[org + 3 * MVG_LT_STRIDE - 1];
            pred = (left && up) ? (sum + 4) >> 3 : (left || up) ? (sum + 2) >> 2 : 128;
        } else {
            const uint32_t taps = lut4[((tr ? 9 : 0) + mode) * 16 + pix];
            const uint8_t *nb = lt + org - MVG_LUT4_BIAS;
            pred = ((int)nb[taps & 255] + (int)nb[(taps >> 8) & 255] + (int)nb[(taps >> 16) & 255] +
                    (int)nb[taps >> 24] + 2) >> 2;
        }
        const int r = s.resid[blk * 16 + pix];
        lt[org + (pxy >> 4) * MVG_LT_STRIDE + (pxy & 3)] = (uint8_t)mvg_add_clip8(pred, r);
    }
    __syncwarp();
}

__device__ __forceinline__ void k2_luma4(K2WarpSmem &s, const uint32_t *lut4, int lane,
                                         unsigned mlo, unsigned mhi, bool availA, bool availB, bool availC)
{
    const int half = lane >> 4, pix = lane & 15;
    const int pxy = ((lane >> 2) & 3) * 16 + (lane & 3);      /* py*16 + px */
    k2_luma4_step<0>(s, lut4, half, pix, pxy, mlo, mhi, availA, availB, availC);
    k2_luma4_step<1>(s, lut4, half, pix, pxy, mlo, mhi, availA, availB, availC);
    k2_luma4_step<2>(s, lut4, half, pix, pxy, mlo, mhi, availA, availB, availC);
    k2_luma4_step<3>(s, lut4, half, pix, pxy, mlo, mhi, availA, availB, availC);
    k2_luma4_step<4>(s, lut4, half, pix, pxy, mlo, mhi, availA, availB, availC);
    k2_luma4_step<5>(s, lut4, half, pix, pxy, mlo, mhi, availA, availB, availC);
    k2_luma4_step<6>(s, lut4, half, pix, pxy, mlo, mhi, availA, availB, availC);
    k2_luma4_step<7>(s, lut4, half, pix, pxy, mlo, mhi, availA, availB, availC);
    k2_luma4_step<8>(s, lut4, half, pix, pxy, mlo, mhi, availA, availB, availC);
    k2_luma4_step<9>(s, lut4, half, pix, pxy, mlo, mhi, availA, availB, availC);
}

/* ---- Intra8x8 luma: 4 blocks in order ------------------------------------------ */
/* The 25 neighbours of a block form one line: n = 0..7 p[-1,7..0], 8 p[-1,-1], 9..24 p[0..15,-1].
 * Lane n filters its entry (reference sample filter, h264_intra_prediction.c:1295-1353) and then
 * derives, again with shuffles, the two smoothings every directional mode is built from:
 *   f2[n] = (p'[n] + p'[n+1] + 1) >> 1,  f3[n] = (p'[n-1] + 2 p'[n] + p'[n+1] + 2) >> 2
 * (line ends replicate).  Each predicted sample is then ONE of p'[i], f2[i], f3[i] -- the
 * table lut8 says which (byte offset of entry i, bit shift of the variant). */
struct K2Lane8 {            /* per-lane constants of the neighbour gather */
    int off_tr, off_notr;   /* tile offset of this lane's neighbour relative to the block origin */
};

template <int B8>
__device__ __forceinline__ void k2_luma8_block(K2WarpSmem &s, const uint16_t *lut8, int lane, const K2Lane8 &k,
                                               unsigned mlo, bool availA, bool availB, bool availC, bool availD)
{
    uint8_t *lt = s.lt;
    constexpr int xo = (B8 & 1) * 8, yo = (B8 >> 1) * 8;
    constexpr int org = (yo + 1) * MVG_LT_STRIDE + MVG_LT_XOFF + xo;
    const int mode = (int)((mlo >> (4 * B8)) & 15);
    const bool left = B8 & 1 ? true : availA, up = B8 & 2 ? true : availB;
    const bool upleft = B8 == 0 ? availD : (B8 == 1 ? availB : (B8 == 2 ? availA : true));
    const bool tr = B8 == 0 ? availB : (B8 == 1 ? availC : (B8 == 2));

    const int raw = lt[org + (tr ? k.off_tr : k.off_notr)];
    int prev = __shfl_up_sync(MVG_FULL, raw, 1), next = __shfl_down_sync(MVG_FULL, raw, 1);
    if (lane == 0) prev = raw;                              /* p'[-1,7] = (p[-1,6] + 3 p[-1,7] + 2) >> 2 */
    if (lane == 24) next = raw;                             /* p'[15,-1] */
    if (!upleft) { if (lane == 7) next = raw; if (lane == 9) prev = raw; }
    if (lane == 8) { if (!left) prev = raw; if (!up) next = raw; }
    const int filt = (prev + 2 * raw + next + 2) >> 2;
    int fp = __shfl_up_sync(MVG_FULL, filt, 1), fn = __shfl_down_sync(MVG_FULL, filt, 1);
    if (lane == 0) fp = filt;
    if (lane == 24) fn = filt;
    const int f2 = (filt + fn + 1) >> 1, f3 = (fp + 2 * filt + fn + 2) >> 2;
    if (lane < 25) s.n8[lane] = (unsigned)filt | ((unsigned)f2 << 8) | ((unsigned)f3 << 16);
    if (mode == 2) {                                        /* warp-uniform */
        int v = 0;
        if (lane < 8 && left) v = filt;
        if (lane >= 9 && lane < 17 && up) v = filt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MVG_FULL, v, o);
        v = (left && up) ? (v + 8) >> 4 : (left || up) ? (v + 4) >> 3 : 128;
        if (lane == 0) s.n8[MVG_N8_DC] = (unsigned)v;
    }
    __syncwarp();
    const int px = lane & 7, py = lane >> 3;
    const uint16_t *lrow = lut8 + mode * 64 + lane;         /* sample (px, py), then (px, py + 4) */
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int y = py + 4 * h;
        const unsigned e = lrow[32 * h];
        const unsigned w = *reinterpret_cast<const unsigned *>(reinterpret_cast<const uint8_t *>(s.n8) + (e & 255));
        const int pred = (w >> (e >> 8)) & 255;
        const int r = s.resid[(B8 * 4 + (y >> 2) * 2 + (px >> 2)) * 16 + (y & 3) * 4 + (px & 3)];
        lt[org + y * MVG_LT_STRIDE + px] = (uint8_t)mvg_add_clip8(pred, r);
    }
    __syncwarp();
}

__device__ __forceinline__ void k2_luma8(K2WarpSmem &s, const uint16_t *lut8, int lane, const K2Lane8 &k,
                                         unsigned mlo, bool availA, bool availB, bool availC, bool availD)
{
    k2_luma8_block<0>(s, lut8, lane, k, mlo, availA, availB, availC, availD);
    k2_luma8_block<1>(s, lut8, lane, k, mlo, availA, availB, availC, availD);
    k2_luma8_block<2>(s, lut8, lane, k, mlo, availA, availB, availC, availD);
    k2_luma8_block<3>(s, lut8, lane, k, mlo, availA, availB, availC, availD);
}

/* ---- chroma, both planes at once (h264_intra_prediction.c:2338-2564) --------- */
__device__ __forceinline__ void k2_chroma(K2WarpSmem &s, int lane, int mode, bool left, bool up)
{
    const int pl = lane >> 4, y = (lane & 15) >> 1, x0 = (lane & 1) * 4;
    uint8_t *ct = s.ct[pl];
    const int top = MVG_CT_XOFF, lcol = MVG_CT_XOFF - 1;
    int pred[4];
    if (mode == 0) {            /* DC, per 4x4 block */
        const int yo = y & 4;
        int st = 0, sl = 0;
        if (up) st = mvg_sum4(*reinterpret_cast<const unsigned *>(ct + top + x0));
        if (left) sl = (int)ct[(yo + 1) * MVG_CT_STRIDE + lcol] + (int)ct[(yo + 2) * MVG_CT_STRIDE + lcol] +
                       (int)ct[(yo + 3) * MVG_CT_STRIDE + lcol] + (int)ct[(yo + 4) * MVG_CT_STRIDE + lcol];
        int v;
        if (!left && !up) v = 128;
        else if ((x0 == 0) == (yo == 0))       /* blocks (0,0) and (4,4) */
            v = (left && up) ? (st + sl + 4) >> 3 : left ? (sl + 2) >> 2 : (st + 2) >> 2;
        else if (x0 > 0) v = up ? (st + 2) >> 2 : (sl + 2) >> 2;          /* (4,0): top first  */
        else v = left ? (sl + 2) >> 2 : (st + 2) >> 2;                    /* (0,4): left first */
        pred[0] = pred[1] = pred[2] = pred[3] = v;
    } else if (mode == 1) {     /* Horizontal */
        const int v = ct[(y + 1) * MVG_CT_STRIDE + lcol];
        pred[0] = pred[1] = pred[2] = pred[3] = v;
    } else if (mode == 2) {     /* Vertical */
        const unsigned t = *reinterpret_cast<const unsigned *>(ct + top + x0);
#pragma unroll
        for (int i = 0; i < 4; i++) pred[i] = (t >> (8 * i)) & 255;
    } else {                    /* Plane */
        int H = 0, V = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            H += (i + 1) * ((int)ct[top + 4 + i] - (int)ct[top + 2 - i]);
            V += (i + 1) * ((int)ct[(5 + i) * MVG_CT_STRIDE + lcol] - (int)ct[(3 - i) * MVG_CT_STRIDE + lcol]);
        }
        const int a = 16 * ((int)ct[8 * MVG_CT_STRIDE + lcol] + (int)ct[top + 7]);
        const int b = (34 * H + 32) >> 6, c = (34 * V + 32) >> 6;
        const int base = a + c * (y - 3) + 16;
#pragma unroll
        for (int k = 0; k < 4; k++) pred[k] = mvg_clip8((base + b * (x0 + k - 3)) >> 5);
    }
    const uint2 r = *reinterpret_cast<const uint2 *>(s.resid + 256 + pl * 64 + ((y >> 2) * 2 + (x0 >> 2)) * 16 + (y & 3) * 4);
    const unsigned p0 = (unsigned)mvg_add_clip8(pred[0], (short)(r.x & 0xffff));
    const unsigned p1 = (unsigned)mvg_add_clip8(pred[1], (int)r.x >> 16);
    const unsigned p2 = (unsigned)mvg_add_clip8(pred[2], (short)(r.y & 0xffff));
    const unsigned p3 = (unsigned)mvg_add_clip8(pred[3], (int)r.y >> 16);
    *reinterpret_cast<unsigned *>(ct + (y + 1) * MVG_CT_STRIDE + MVG_CT_XOFF + x0) = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
}

/* Persistent warps.  A work item is one macroblock row of one picture; a warp claims items
 * from an atomic counter and walks its row left to right.
 *
 * Wavefront dependency: row r may process macroblock x once row r-1 has finished macroblock
 * x+1 (its up-right neighbour C, h264_spatial.c:371-382).  The only samples that cross rows are
 * the bottom sample line of the macroblocks above, so a finished macroblock publishes that
 * line (16 Y + 8 Cb + 8 Cr bytes) as eight 64-bit words {4 data bytes, launch epoch} -- the
 * flag travels with the data, every word validates itself, and neither side needs a fence or
 * a separate progress counter.  The row below spins only on words whose epoch is stale.
 *
 * Claim order: pictures are taken in groups of `group`; inside a group items are ordered
 * row-major over (row, picture).  Row r-1 of a picture is therefore always claimed before row
 * r (no deadlock: it runs on a resident warp), and `group` items earlier, so in steady state
 * it is many macroblocks ahead and the spin is rarely entered. */
__global__ void __launch_bounds__(K2_WARPS * 32, 8)
k2_wavefront(K2Params p)
{
    __shared__ uint32_t s_lut4[2 * 9 * 16];
    __shared__ uint16_t s_lut8[9 * 64];
    __shared__ K2WarpSmem s_warp[K2_WARPS];

    for (int i = threadIdx.x; i < 2 * 9 * 16; i += blockDim.x) s_lut4[i] = (&p.luts->lut4[0][0][0])[i];
    for (int i = threadIdx.x; i < 9 * 64 / 2; i += blockDim.x)
        reinterpret_cast<uint32_t *>(s_lut8)[i] = reinterpret_cast<const uint32_t *>(&p.luts->lut8[0][0])[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    K2WarpSmem &s = s_warp[threadIdx.x >> 5];
    const int W = p.w_mbs, H = p.h_mbs, n_mb = W * H;
    const int ystride = W * 16, cstride = W * 8;
    const size_t pic_bytes = (size_t)n_mb * 384;
    const int total = p.n_pics * H;
    const unsigned epoch = p.epoch;

    /* per-lane constants --------------------------------------------------------------- */
    /* halo word carried by lanes 0..7: 0..3 luma x = 4*lane, 4,5 Cb, 6,7 Cr */
    uint8_t *const halo_top = lane < 4 ? s.lt + MVG_LT_XOFF + lane * 4
                                       : s.ct[(lane >> 1) & 1] + MVG_CT_XOFF + (lane & 1) * 4;      /* sample row -1 */
    const uint8_t *const halo_bot = lane < 4 ? halo_top + 16 * MVG_LT_STRIDE : halo_top + 8 * MVG_CT_STRIDE;
    /* column x = 15 -> x = -1 hand-over: lanes 0..16 luma rows -1..15, lanes 17..25 Cb rows -1..7, and a
     * second move by lanes 0..8 for Cr */
    uint8_t *const lc_dst = lane < 17 ? s.lt + lane * MVG_LT_STRIDE + MVG_LT_XOFF - 1
                                      : s.ct[0] + (lane < 26 ? lane - 17 : 0) * MVG_CT_STRIDE + MVG_CT_XOFF - 1;
    const int lc_span = lane < 17 ? 16 : 8;
    uint8_t *const lc_dst2 = s.ct[1] + (lane < 9 ? lane : 0) * MVG_CT_STRIDE + MVG_CT_XOFF - 1;
    /* picture write-out: lanes 0..15 one luma row (16 B), lanes 16..23 Cb rows, 24..31 Cr rows (8 B) */
    const uint8_t *const wo_src = lane < 16 ? s.lt + (lane + 1) * MVG_LT_STRIDE + MVG_LT_XOFF
                                            : s.ct[(lane >> 3) & 1] + ((lane & 7) + 1) * MVG_CT_STRIDE + MVG_CT_XOFF;
    /* Intra8x8 neighbour gather */
    K2Lane8 k8;
    if (lane < 8)       k8.off_tr = (7 - lane) * MVG_LT_STRIDE - 1;
    else if (lane == 8) k8.off_tr = -MVG_LT_STRIDE - 1;
    else if (lane < 25) k8.off_tr = -MVG_LT_STRIDE + (lane - 9);
    else                k8.off_tr = 0;
    k8.off_notr = (lane > 16 && lane < 25) ? -MVG_LT_STRIDE + 7 : k8.off_tr;     /* p[8..15,-1] := p[7,-1] */

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(p.work, 1);
        item = __shfl_sync(MVG_FULL, item, 0);
        if (item >= total) break;
        const int g = item / (p.group * H);
        const int gsize = min(p.group, p.n_pics - g * p.group);
        const int within = item - g * p.group * H;
        const int row = within / gsize;
        const int slot = p.first_slot + g * p.group + (within - row * gsize);

        uint8_t *ybase = p.yuv + (size_t)slot * pic_bytes;
        /* where this lane writes its row of every macroblock of this macroblock row */
        uint8_t *wo_dst = lane < 16 ? ybase + (size_t)(row * 16 + lane) * ystride
                                    : ybase + (size_t)n_mb * (lane < 24 ? 256 : 320) + (size_t)(row * 8 + (lane & 7)) * cstride;
        const int wo_step = lane < 16 ? 16 : 8;
        const int16_t *resid = p.resid + ((size_t)slot * n_mb + (size_t)row * W) * 384;
        const MvgMbCtl *ctl = p.ctl + (size_t)slot * n_mb + (size_t)row * W;
        const bool availB = row > 0, publish = row < H - 1;
        const uint2 *habove = p.halo + ((size_t)slot * n_mb + (size_t)(row - 1) * W) * 8 + lane;
        uint2 *hmine = p.halo + ((size_t)slot * n_mb + (size_t)row * W) * 8 + lane;

        /* prefetch the first macroblock's inputs */
        uint4 r0 = __ldg(reinterpret_cast<const uint4 *>(resid) + lane);
        uint4 r1 = make_uint4(0, 0, 0, 0);
        if (lane < 16) r1 = __ldg(reinterpret_cast<const uint4 *>(resid) + 32 + lane);
        uint4 c4 = __ldg(reinterpret_cast<const uint4 *>(ctl));
        /* halo words of the row above: cur = macroblock mx, nxt = macroblock mx+1 */
        uint2 cur = make_uint2(0, 0), nxt = make_uint2(0, 0);
        if (availB && lane < 8) {
            cur = mvg_ld_relaxed_u64(habove);
            if (W > 1) nxt = mvg_ld_relaxed_u64(habove + 8);
        }
        if (availB) {
            unsigned ns = 64;
            while (!__all_sync(MVG_FULL, lane >= 8 || cur.y == epoch)) {
                __nanosleep(ns);
                if (ns < 2048) ns *= 2;
                if (lane < 8) cur = mvg_ld_relaxed_u64(habove);
            }
        }

        for (int mx = 0; mx < W; mx++) {
            reinterpret_cast<uint4 *>(s.resid)[lane] = r0;
            if (lane < 16) reinterpret_cast<uint4 *>(s.resid)[32 + lane] = r1;
            const uint4 ctlw = c4;
            if (mx + 1 < W) {       /* software pipeline: next macroblock's loads fly during this one */
                const uint4 *nr = reinterpret_cast<const uint4 *>(resid + (size_t)(mx + 1) * 384);
                r0 = __ldg(nr + lane);
                if (lane < 16) r1 = __ldg(nr + 32 + lane);
                c4 = __ldg(reinterpret_cast<const uint4 *>(ctl + mx + 1));
            }
            const bool availA = mx > 0, availC = availB && mx < W - 1, availD = availA && availB;

            if (availB) {
                if (availC) {       /* the up-right macroblock must have been published */
                    /* rarely entered; when it is, this row has caught up with the row above, so sleep
                     * for about a macroblock time instead of polling at full rate */
                    unsigned ns = 400;
                    while (!__all_sync(MVG_FULL, lane >= 8 || nxt.y == epoch)) {
                        __nanosleep(ns);
                        if (ns < 3200) ns *= 2;
                        if (lane < 8) nxt = mvg_ld_relaxed_u64(habove + (size_t)(mx + 1) * 8);
                    }
                }
                /* sample row -1 of the tiles: x = 0..15 from cur, x = 16..23 from nxt */
                if (lane < 8) *reinterpret_cast<unsigned *>(halo_top) = cur.x;
                if (lane < 2) *reinterpret_cast<unsigned *>(halo_top + 16) = nxt.x;
                cur = nxt;
                if (mx + 2 < W && lane < 8) nxt = mvg_ld_relaxed_u64(habove + (size_t)(mx + 2) * 8);
            }
            __syncwarp();

            const int kind = ctlw.x & 255, i16 = (ctlw.x >> 8) & 255, cmode = (ctlw.x >> 16) & 255;
            if (kind == MVG_MB_I16x16)    k2_luma16(s, lane, i16, availA, availB);
            else if (kind == MVG_MB_I4x4) k2_luma4(s, s_lut4, lane, ctlw.y, ctlw.z, availA, availB, availC);
            else                          k2_luma8(s, s_lut8, lane, k8, ctlw.y, availA, availB, availC, availD);
            k2_chroma(s, lane, cmode, availA, availB);
            __syncwarp();

            /* write the macroblock to the planar picture */
            if (lane < 16) {
                const uint2 a = *reinterpret_cast<const uint2 *>(wo_src);
                const uint2 b = *reinterpret_cast<const uint2 *>(wo_src + 8);
                *reinterpret_cast<uint4 *>(wo_dst) = make_uint4(a.x, a.y, b.x, b.y);
            } else {
                *reinterpret_cast<uint2 *>(wo_dst) = *reinterpret_cast<const uint2 *>(wo_src);
            }
            wo_dst += wo_step;
            /* publish the bottom sample line for the row below */
            if (publish && lane < 8)
                mvg_st_relaxed_u64(hmine + (size_t)mx * 8, *reinterpret_cast<const unsigned *>(halo_bot), epoch);
            /* next macroblock: x = 15 becomes x = -1 (luma rows -1..15, chroma x = 7, rows -1..7) */
            if (lane < 26) *lc_dst = lc_dst[lc_span];
            if (lane < 9) *lc_dst2 = lc_dst2[8];
            __syncwarp();
        }
    }
}

/* ========================================================================= */
/* Kernel 3                                                                    */

struct K3Params {
    const uint8_t *yuv;     /* [slot][1.5*W*H] */
    uint8_t       *rgb;     /* [slot][3*(W/s)*(H/s)] */
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

/* scale 1: one thread converts a 16 x 2 sample patch: 2 x 16 B of Y, 8 B of Cb and of Cr in
 * (each chroma sample covers a 2 x 2 patch, export_utils.c:278-279), 2 x 48 B of RGB24 out as
 * 128-bit stores.  The chroma contributions are computed once per chroma sample. */
__global__ void __launch_bounds__(256)
k3_rgb_full(K3Params p)
{
    const int groups_per_row = p.width >> 4, row_pairs = p.height >> 1;
    const long long per_pic = (long long)groups_per_row * row_pairs;
    const long long total = per_pic * p.n_pics;
    const size_t ysz = (size_t)p.width * p.height;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int pic = (int)(g / per_pic);
        const int rem = (int)(g - (long long)pic * per_pic);
        const int yp = rem / groups_per_row, gx = rem - yp * groups_per_row;
        const size_t slot = (size_t)(p.first_slot + pic);
        const uint8_t *Y = p.yuv + slot * (ysz * 3 / 2);
        const uint8_t *Cb = Y + ysz, *Cr = Cb + ysz / 4;
        const uint8_t *yrow = Y + (size_t)(2 * yp) * p.width + gx * 16;
        const uint4 y0 = __ldg(reinterpret_cast<const uint4 *>(yrow));
        const uint4 y1 = __ldg(reinterpret_cast<const uint4 *>(yrow + p.width));
        const size_t coff = (size_t)yp * (p.width >> 1) + gx * 8;
        const uint2 cb = __ldg(reinterpret_cast<const uint2 *>(Cb + coff));
        const uint2 cr = __ldg(reinterpret_cast<const uint2 *>(Cr + coff));
        const unsigned yw[2][4] = {{y0.x, y0.y, y0.z, y0.w}, {y1.x, y1.y, y1.z, y1.w}};
        const unsigned cbw[2] = {cb.x, cb.y}, crw[2] = {cr.x, cr.y};
        unsigned out[2][12];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int k = 0; k < 12; k++) out[r][k] = 0;
#pragma unroll
        for (int ci = 0; ci < 8; ci++) {
            const int Cbv = (cbw[ci >> 2] >> (8 * (ci & 3))) & 255, Crv = (crw[ci >> 2] >> (8 * (ci & 3))) & 255;
            /* export_utils.c:300-302, the terms that do not depend on Y */
            const int rC = ((408 * Crv) >> 8) - 222;
            const int gC = 135 - ((100 * Cbv) >> 8) - ((208 * Crv) >> 8);
            const int bC = ((516 * Cbv) >> 8) - 276;
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int i = 2 * ci + h;
                    const int Yv = (yw[r][i >> 2] >> (8 * (i & 3))) & 255;
                    const int t = (298 * Yv) >> 8;
                    const unsigned R = (unsigned)__viaddmin_s32_relu(t, rC, 255);
                    const unsigned G = (unsigned)__viaddmin_s32_relu(t, gC, 255);
                    const unsigned B = (unsigned)__viaddmin_s32_relu(t, bC, 255);
                    const int b0 = 3 * i, b1 = 3 * i + 1, b2 = 3 * i + 2;
                    out[r][b0 >> 2] |= R << (8 * (b0 & 3));
                    out[r][b1 >> 2] |= G << (8 * (b1 & 3));
                    out[r][b2 >> 2] |= B << (8 * (b2 & 3));
                }
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
            uint4 *dst = reinterpret_cast<uint4 *>(p.rgb + slot * (ysz * 3) + ((size_t)(2 * yp + r) * p.width + gx * 16) * 3);
            dst[0] = make_uint4(out[r][0], out[r][1], out[r][2], out[r][3]);
            dst[1] = make_uint4(out[r][4], out[r][5], out[r][6], out[r][7]);
            dst[2] = make_uint4(out[r][8], out[r][9], out[r][10], out[r][11]);
        }
    }
}

/* scale s > 1: one thread per output pixel, rounded s x s box average of the
 * full-resolution RGB picture. */
__global__ void __launch_bounds__(256)
k3_rgb_scaled(K3Params p)
{
    const int s = p.scale, ow = p.width / s, oh = p.height / s;
    const long long per_pic = (long long)ow * oh, total = per_pic * p.n_pics;
    const size_t ysz = (size_t)p.width * p.height;
    const int area = s * s;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int pic = (int)(g / per_pic);
        const long long rem = g - (long long)pic * per_pic;
        const int oy = (int)(rem / ow), ox = (int)(rem - (long long)oy * ow);
        const size_t slot = (size_t)(p.first_slot + pic);
        const uint8_t *Y = p.yuv + slot * (ysz * 3 / 2);
        const uint8_t *Cb = Y + ysz, *Cr = Cb + ysz / 4;
        int aR = 0, aG = 0, aB = 0;
        for (int dy = 0; dy < s; dy++) {
            const int py = oy * s + dy;
            for (int dx = 0; dx < s; dx++) {
                const int px = ox * s + dx;
                const size_t co = (size_t)(py >> 1) * (p.width >> 1) + (px >> 1);
                int R, G, B;
                mvg_ycc_to_rgb(__ldg(Y + (size_t)py * p.width + px), __ldg(Cb + co), __ldg(Cr + co), R, G, B);
                aR += R; aG += G; aB += B;
            }
        }
        uint8_t *o = p.rgb + slot * ((size_t)ow * oh * 3) + ((size_t)oy * ow + ox) * 3;
        o[0] = (uint8_t)((aR + area / 2) / area);
        o[1] = (uint8_t)((aG + area / 2) / area);
        o[2] = (uint8_t)((aB + area / 2) / area);
    }
}
