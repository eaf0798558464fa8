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

#define K1_WARPS 8

/* One warp per macroblock, grid-stride, the next macroblock's loads in flight while the
 * current one is transformed.
 *  4x4 path (Intra4x4 / Intra16x16 luma, all chroma): lane b < 24 owns 4x4 block b
 *    (0..15 luma in decoding order, 16..19 Cb, 20..23 Cr): two 128-bit loads bring
 *    its 16 levels, the zig-zag inverse is a compile-time register renaming, both
 *    butterfly passes stay in registers.  The Intra16x16 DC Hadamard and the chroma
 *    DC 2x2 run across lanes with shuffles.
 *  8x8 path (Intra8x8 luma): 8 lanes per block, one matrix row per lane, transposed
 *    through shared memory between the row and the column pass.
 *  The dequantisation scale (LevelScale << (qP/6 - 4)) is tabulated per qP in shared
 *  memory; the rounding constant 32 of the final >> 6 is added to the DC term before the
 *  butterflies (it reaches every output with weight one).  The int16 residual is staged
 *  in shared memory in raster order and leaves with coalesced 128-bit stores. */
__global__ void __launch_bounds__(K1_WARPS * 32, 3)
k1_dequant_idct(K1Params p)
{
    __shared__ int32_t s_ls4[3 * 6 * 16];
    __shared__ int32_t s_ls4q[3 * 52 * 16];                     /* qP >= 24: LevelScale << (qP/6-4) */
    __shared__ int32_t s_ls8[6 * 64];
    __shared__ uint8_t s_zz8inv[64];
    __shared__ __align__(16) int16_t s_in[K1_WARPS][256];      /* luma levels of an Intra8x8 MB */
    __shared__ int32_t s_tr[K1_WARPS][4][8][9];                 /* 8x8 transpose, padded          */
    __shared__ __align__(16) int16_t s_out[K1_WARPS][384];     /* residual, raster                */

    for (int i = threadIdx.x; i < 3 * 6 * 16; i += blockDim.x) s_ls4[i] = (&p.tab->ls4[0][0][0])[i];
    for (int i = threadIdx.x; i < 3 * 52 * 16; i += blockDim.x) s_ls4q[i] = (&p.tab->ls4q[0][0][0])[i];
    for (int i = threadIdx.x; i < 6 * 64; i += blockDim.x) s_ls8[i] = (&p.tab->ls8[0][0])[i];
    if (threadIdx.x < 64) s_zz8inv[threadIdx.x] = p.tab->zz8inv[threadIdx.x];
    const int cb_off = p.tab->cb_qp_offset, cr_off = p.tab->cr_qp_offset;
    __syncthreads();

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long n_warps = (long long)gridDim.x * K1_WARPS;
    const bool is_chroma = lane >= 16;
    const int comp = lane < 16 ? 0 : (lane < 20 ? 1 : 2);
    int16_t *out = s_out[w];
    /* where this lane's 4x4 block lands in the raster residual */
    int16_t *dst;
    int dst_stride;
    if (lane < 16) {
        const int bx = (lane & 1) | (((lane >> 2) & 1) << 1), by = ((lane >> 1) & 1) | ((lane >> 3) << 1);
        dst = out + (by * 4) * 16 + bx * 4; dst_stride = 16;
    } else {
        const int b = lane & 3, pl = (lane >> 2) & 1;
        dst = out + 256 + pl * 64 + ((b >> 1) * 4) * 8 + (b & 1) * 4; dst_stride = 8;
    }

    /* per-MB side information, one byte per lane: lanes 0..15 luma modes, 16 mb_kind,
     * 17 Intra16x16PredMode, 18 intra_chroma_pred_mode, 19 QPY */
    const uint8_t *meta_src = lane < 16 ? p.luma_modes + lane
                            : lane == 16 ? p.mb_kind : lane == 17 ? p.i16_mode
                            : lane == 18 ? p.chroma_mode : reinterpret_cast<const uint8_t *>(p.qp_y);
    const long long meta_stride = lane < 16 ? 16 : 1;
    long long mb = (long long)blockIdx.x * K1_WARPS + w;
    uint4 na = make_uint4(0, 0, 0, 0), nb = na;
    int nmeta = 0;
    if (mb < p.n_mbs) {
        if (lane < 24) {
            const uint4 *src = reinterpret_cast<const uint4 *>(p.coeff + mb * 384 + lane * 16);
            na = __ldg(src); nb = __ldg(src + 1);
        }
        if (lane < 20) nmeta = __ldg(meta_src + mb * meta_stride);
    }

    for (; mb < p.n_mbs; mb += n_warps) {
        const uint4 a = na, b = nb;
        const int mode_byte = nmeta;
        const int kind = __shfl_sync(MVG_FULL, nmeta, 16);
        const int qp = (signed char)__shfl_sync(MVG_FULL, nmeta, 19);
        const unsigned w0 = (unsigned)kind | ((unsigned)__shfl_sync(MVG_FULL, nmeta, 17) << 8) |
                            ((unsigned)__shfl_sync(MVG_FULL, nmeta, 18) << 16);
        {   /* prefetch the next macroblock of this warp */
            const long long nx = mb + n_warps;
            if (nx < p.n_mbs) {
                if (lane < 24) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(p.coeff + nx * 384 + lane * 16);
                    na = __ldg(src); nb = __ldg(src + 1);
                }
                if (lane < 20) nmeta = __ldg(meta_src + nx * meta_stride);
            }
        }

        /* ---------------- 4x4 blocks ---------------- */
        const bool luma4 = (kind != MVG_MB_I8x8);
        const bool active = lane < 24 && (is_chroma || luma4);
        int c[16];                              /* matrix, row-major, after inverse zig-zag */
        {
            /* zig-zag k -> (row,col): utils.h:64 / spec Table 8-13 */
            c[0] = (short)(a.x & 0xffff); c[1] = (int)a.x >> 16; c[4] = (short)(a.y & 0xffff); c[8] = (int)a.y >> 16;
            c[5] = (short)(a.z & 0xffff); c[2] = (int)a.z >> 16; c[3] = (short)(a.w & 0xffff); c[6] = (int)a.w >> 16;
            c[9] = (short)(b.x & 0xffff); c[12] = (int)b.x >> 16; c[13] = (short)(b.y & 0xffff); c[10] = (int)b.y >> 16;
            c[7] = (short)(b.z & 0xffff); c[11] = (int)b.z >> 16; c[14] = (short)(b.w & 0xffff); c[15] = (int)b.w >> 16;
        }
        if (kind == MVG_MB_I8x8) {              /* warp-uniform: stage the luma levels for the 8x8 path */
            if (lane < 16) {
                reinterpret_cast<uint4 *>(s_in[w])[2 * lane] = a;
                reinterpret_cast<uint4 *>(s_in[w])[2 * lane + 1] = b;
            }
        }

        /* quantiser of this lane's block */
        int qpb = qp;
        if (comp) qpb = mvg_chroma_qp(qp, comp == 1 ? cb_off : cr_off);
        const int qd = qpb / 6, qm = qpb - 6 * qd;
        const int ls00 = s_ls4[(comp * 6 + qm) * 16];
        bool keep_dc = is_chroma;

        /* Intra16x16 luma DC: f = H c H over the 16 lanes (h264_transform.c:783-808) */
        if (kind == MVG_MB_I16x16) {            /* warp-uniform */
            const int bi = ((lane >> 1) & 1) | (((lane >> 3) & 1) << 1);
            const int bj = (lane & 1) | (((lane >> 2) & 1) << 1);
            /* rows of H (h264_transform.c:62-68) as sign masks over k: 0000, 1100, 0110, 1010 */
            const unsigned hs_i = (0xA6C0u >> (4 * bi)) & 0xF;
            const unsigned hs_j = (0xA6C0u >> (4 * bj)) & 0xF;
            int f = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int l = 0; l < 4; l++) {
                    const int src = (l & 1) | ((k & 1) << 1) | ((l >> 1) << 2) | ((k >> 1) << 3);
                    const int v = __shfl_sync(MVG_FULL, c[0], src);
                    const unsigned neg = ((hs_i >> k) ^ (hs_j >> l)) & 1u;
                    f += neg ? -v : v;
                }
            const int t = f * ls00;
            const int dcy = (qp >= 36) ? (int)((unsigned)t << (qd - 6)) : ((t + (1 << (5 - qd))) >> (6 - qd));
            if (lane < 16) { c[0] = dcy; keep_dc = true; }
        }
        /* chroma DC 2x2 (h264_transform.c:988-1005, :924-936) over lane groups 16..19, 20..23 */
        {
            const int base = lane & ~3, r = (lane >> 1) & 1, cc = lane & 1;
            int f = 0;
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const int v = __shfl_sync(MVG_FULL, c[0], base + s);
                const int neg = (r & (s >> 1)) ^ (cc & (s & 1));
                f += neg ? -v : v;
            }
            if (is_chroma && lane < 24) c[0] = (int)((unsigned)(f * ls00) << qd) >> 5;
        }

        unsigned nz_any = 0;
        if (active) {
#pragma unroll
            for (int k = 0; k < 16; k++) nz_any |= (unsigned)c[k];
            /* quant4x4, h264_transform.c:1100-1134 */
            const int dc_in = c[0];
            const int32_t *lq = s_ls4q + (comp * 52 + qpb) * 16;
            if (qpb > 23) {
#pragma unroll
                for (int k = 0; k < 16; k++) c[k] = c[k] * lq[k];
            } else {
                const int rnd = 1 << (3 - qd), sh = 4 - qd;
#pragma unroll
                for (int k = 0; k < 16; k++) c[k] = (c[k] * lq[k] + rnd) >> sh;
            }
            if (keep_dc) c[0] = dc_in;
            c[0] += 32;                         /* rounding of the final >> 6 (h264_transform.c:1190) */
            /* idct4x4: rows then columns */
#pragma unroll
            for (int i = 0; i < 4; i++)
                mvg_bfly4(c[4 * i], c[4 * i + 1], c[4 * i + 2], c[4 * i + 3], c[4 * i], c[4 * i + 1], c[4 * i + 2], c[4 * i + 3]);
#pragma unroll
            for (int j = 0; j < 4; j++)
                mvg_bfly4(c[j], c[4 + j], c[8 + j], c[12 + j], c[j], c[4 + j], c[8 + j], c[12 + j]);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint2 pk;
                pk.x = mvg_pack_sat16(c[4 * i + 1] >> 6, c[4 * i] >> 6);
                pk.y = mvg_pack_sat16(c[4 * i + 3] >> 6, c[4 * i + 2] >> 6);
                *reinterpret_cast<uint2 *>(dst + i * dst_stride) = pk;
            }
        }
        unsigned nzmask = __ballot_sync(MVG_FULL, nz_any != 0) & (luma4 ? 0x00FFFFFFu : 0x00FF0000u);

        /* ---------------- Intra8x8 luma ---------------- */
        if (kind == MVG_MB_I8x8) {              /* warp-uniform */
            __syncwarp();
            const int b8 = lane >> 3, row = lane & 7;
            const int qd8 = qp / 6, qm8 = qp - 6 * qd8;
            const int32_t *l8 = s_ls8 + qm8 * 64 + row * 8;
            int v[8];
            unsigned any8 = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int lvl = s_in[w][b8 * 64 + s_zz8inv[row * 8 + j]];
                any8 |= (unsigned)lvl;
                const int t = lvl * l8[j];                /* quant8x8, h264_transform.c:1256-1284 */
                v[j] = (qp > 35) ? (int)((unsigned)t << (qd8 - 6)) : ((t + (1 << (5 - qd8))) >> (6 - qd8));
            }
            if (row == 0) v[0] += 32;                     /* rounding of the final >> 6 (:1382) */
            mvg_idct8_1d(v);                              /* row pass */
#pragma unroll
            for (int j = 0; j < 8; j++) s_tr[w][b8][row][j] = v[j];
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = s_tr[w][b8][i][row];   /* this lane now owns column `row` */
            mvg_idct8_1d(v);                              /* column pass */
            const int xo = (b8 & 1) * 8 + row, yo = (b8 >> 1) * 8;
#pragma unroll
            for (int i = 0; i < 8; i++) out[(yo + i) * 16 + xo] = (int16_t)min(max(v[i] >> 6, -32768), 32767);
            const unsigned bal = __ballot_sync(MVG_FULL, any8 != 0);
#pragma unroll
            for (int q = 0; q < 4; q++) if ((bal >> (8 * q)) & 0xFFu) nzmask |= 0xFu << (4 * q);
        }
        __syncwarp();

        /* ---------------- write residual + control record ---------------- */
        uint4 *gout = reinterpret_cast<uint4 *>(p.resid + mb * 384);
        const uint4 *sout = reinterpret_cast<const uint4 *>(out);
        gout[lane] = sout[lane];
        if (lane < 16) gout[32 + lane] = sout[32 + lane];

        /* 16 prediction modes -> 16 nibbles: OR-reduce inside each group of 8 lanes */
        unsigned nib = lane < 16 ? (unsigned)(mode_byte & 15) << (4 * (lane & 7)) : 0u;
        nib |= __shfl_xor_sync(MVG_FULL, nib, 1);
        nib |= __shfl_xor_sync(MVG_FULL, nib, 2);
        nib |= __shfl_xor_sync(MVG_FULL, nib, 4);
        const unsigned w2 = __shfl_sync(MVG_FULL, nib, 8);
        if (lane == 0) *reinterpret_cast<uint4 *>(p.ctl + mb) = make_uint4(w0, nib, w2, nzmask);
        __syncwarp();
    }
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
    __align__(16) uint8_t  n8[32];
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
    const uint4 r = *reinterpret_cast<const uint4 *>(s.resid + y * 16 + x0);
    const unsigned rr[4] = {r.x, r.y, r.z, r.w};
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
        const int r = s.resid[by * 64 + bx * 4 + pxy];
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

/* ---- Intra8x8 luma: 4 blocks in order, reference sample filter + taps -------- */
__device__ __forceinline__ void k2_luma8(K2WarpSmem &s, const uint32_t *lut8, int lane,
                                         unsigned mlo, bool availA, bool availB, bool availC, bool availD)
{
    uint8_t *lt = s.lt;
#pragma unroll 1
    for (int b8 = 0; b8 < 4; b8++) {
        const int xo = (b8 & 1) * 8, yo = (b8 >> 1) * 8;
        const int mode = (int)((mlo >> (4 * b8)) & 15);
        const bool left = xo > 0 || availA, up = yo > 0 || availB;
        const bool upleft = b8 == 0 ? availD : (b8 == 1 ? availB : (b8 == 2 ? availA : true));
        const bool tr = b8 == 0 ? availB : (b8 == 1 ? availC : (b8 == 2));
        const int org = (yo + 1) * MVG_LT_STRIDE + MVG_LT_XOFF + xo;

        /* neighbour line: n = 0..7 left[7..0], 8 corner, 9..24 top[0..15] */
        int raw = 0;
        if (lane < 8)        raw = lt[org + (7 - lane) * MVG_LT_STRIDE - 1];
        else if (lane == 8)  raw = lt[org - MVG_LT_STRIDE - 1];
        else if (lane < 25) {
            int i = lane - 9;
            if (i > 7 && !tr) i = 7;             /* h264_intra_prediction.c:1230-1236 */
            raw = lt[org - MVG_LT_STRIDE + i];
        }
        int prev = __shfl_up_sync(MVG_FULL, raw, 1), next = __shfl_down_sync(MVG_FULL, raw, 1);
        /* h264_intra_prediction.c:1295-1353 */
        if (lane == 0) prev = raw;                              /* p'[-1,7] = (p[-1,6] + 3 p[-1,7] + 2) >> 2 */
        if (lane == 24) next = raw;                             /* p'[15,-1] */
        if (lane == 7 && !upleft) next = raw;                   /* p'[-1,0] without corner */
        if (lane == 9 && !upleft) prev = raw;                   /* p'[0,-1] without corner */
        if (lane == 8) { if (!left) prev = raw; if (!up) next = raw; }
        const int filt = (prev + 2 * raw + next + 2) >> 2;
        if (lane < 25) s.n8[lane] = (uint8_t)filt;
        if (mode == 2) {                                        /* warp-uniform */
            int v = 0;
            if (lane < 8 && left) v = filt;
            if (lane >= 9 && lane < 17 && up) v = filt;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MVG_FULL, v, o);
            v = (left && up) ? (v + 8) >> 4 : (left || up) ? (v + 4) >> 3 : 128;
            if (lane == 0) s.n8[MVG_N8_DC] = (uint8_t)v;
        }
        __syncwarp();
        const int px = lane & 7, py = lane >> 3;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int y = py + 4 * h;
            const uint32_t taps = lut8[mode * 64 + y * 8 + px];
            const int pred = ((int)s.n8[taps & 255] + (int)s.n8[(taps >> 8) & 255] +
                              (int)s.n8[(taps >> 16) & 255] + (int)s.n8[taps >> 24] + 2) >> 2;
            const int r = s.resid[(yo + y) * 16 + xo + px];
            lt[org + y * MVG_LT_STRIDE + px] = (uint8_t)mvg_add_clip8(pred, r);
        }
        __syncwarp();
    }
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
    const uint2 r = *reinterpret_cast<const uint2 *>(s.resid + 256 + pl * 64 + y * 8 + x0);
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
    __shared__ uint32_t s_lut8[9 * 64];
    __shared__ K2WarpSmem s_warp[K2_WARPS];

    for (int i = threadIdx.x; i < 2 * 9 * 16; i += blockDim.x) s_lut4[i] = (&p.luts->lut4[0][0][0])[i];
    for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) s_lut8[i] = (&p.luts->lut8[0][0])[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    K2WarpSmem &s = s_warp[threadIdx.x >> 5];
    const int W = p.w_mbs, H = p.h_mbs, n_mb = W * H;
    const int ystride = W * 16, cstride = W * 8;
    const size_t pic_bytes = (size_t)n_mb * 384;
    const int total = p.n_pics * H;
    const unsigned epoch = p.epoch;

    /* byte offset inside the tiles of the halo word this lane carries (lanes 0..7) */
    const int halo_pl = lane < 4 ? -1 : ((lane - 4) >> 1);
    const int halo_x = lane < 4 ? lane * 4 : ((lane - 4) & 1) * 4;

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
        uint8_t *cbbase = ybase + (size_t)n_mb * 256, *crbase = cbbase + (size_t)n_mb * 64;
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
                    unsigned ns = 64;
                    while (!__all_sync(MVG_FULL, lane >= 8 || nxt.y == epoch)) {
                        __nanosleep(ns);
                        if (ns < 2048) ns *= 2;
                        if (lane < 8) nxt = mvg_ld_relaxed_u64(habove + (size_t)(mx + 1) * 8);
                    }
                }
                /* sample row -1 of the tiles: x = 0..15 from cur, x = 16..23 from nxt */
                if (lane < 4) {
                    *reinterpret_cast<unsigned *>(s.lt + MVG_LT_XOFF + halo_x) = cur.x;
                    if (lane < 2) *reinterpret_cast<unsigned *>(s.lt + MVG_LT_XOFF + 16 + halo_x) = nxt.x;
                } else if (lane < 8) {
                    *reinterpret_cast<unsigned *>(s.ct[halo_pl] + MVG_CT_XOFF + halo_x) = cur.x;
                }
                cur = nxt;
                if (mx + 2 < W && lane < 8) nxt = mvg_ld_relaxed_u64(habove + (size_t)(mx + 2) * 8);
            }
            __syncwarp();

            const int kind = ctlw.x & 255, i16 = (ctlw.x >> 8) & 255, cmode = (ctlw.x >> 16) & 255;
            if (kind == MVG_MB_I16x16)    k2_luma16(s, lane, i16, availA, availB);
            else if (kind == MVG_MB_I4x4) k2_luma4(s, s_lut4, lane, ctlw.y, ctlw.z, availA, availB, availC);
            else                          k2_luma8(s, s_lut8, lane, ctlw.y, availA, availB, availC, availD);
            k2_chroma(s, lane, cmode, availA, availB);
            __syncwarp();

            /* write the macroblock to the planar picture */
            if (lane < 16) {
                const uint2 a = *reinterpret_cast<const uint2 *>(s.lt + (lane + 1) * MVG_LT_STRIDE + MVG_LT_XOFF);
                const uint2 b = *reinterpret_cast<const uint2 *>(s.lt + (lane + 1) * MVG_LT_STRIDE + MVG_LT_XOFF + 8);
                *reinterpret_cast<uint4 *>(ybase + (size_t)(row * 16 + lane) * ystride + mx * 16) = make_uint4(a.x, a.y, b.x, b.y);
            } else {
                const int pl = (lane >> 3) & 1, y = lane & 7;
                const uint2 a = *reinterpret_cast<const uint2 *>(s.ct[pl] + (y + 1) * MVG_CT_STRIDE + MVG_CT_XOFF);
                *reinterpret_cast<uint2 *>((pl ? crbase : cbbase) + (size_t)(row * 8 + y) * cstride + mx * 8) = a;
            }
            /* publish the bottom sample line for the row below */
            if (publish && lane < 8) {
                const unsigned d = lane < 4
                    ? *reinterpret_cast<const unsigned *>(s.lt + 16 * MVG_LT_STRIDE + MVG_LT_XOFF + halo_x)
                    : *reinterpret_cast<const unsigned *>(s.ct[halo_pl] + 8 * MVG_CT_STRIDE + MVG_CT_XOFF + halo_x);
                mvg_st_relaxed_u64(hmine + (size_t)mx * 8, d, epoch);
            }
            /* next macroblock: x = 15 becomes x = -1 (luma rows -1..15, chroma x = 7, rows -1..7) */
            if (lane < 17) s.lt[lane * MVG_LT_STRIDE + MVG_LT_XOFF - 1] = s.lt[lane * MVG_LT_STRIDE + MVG_LT_XOFF + 15];
            else if (lane < 26) s.ct[0][(lane - 17) * MVG_CT_STRIDE + MVG_CT_XOFF - 1] = s.ct[0][(lane - 17) * MVG_CT_STRIDE + MVG_CT_XOFF + 7];
            if (lane < 9) s.ct[1][lane * MVG_CT_STRIDE + MVG_CT_XOFF - 1] = s.ct[1][lane * MVG_CT_STRIDE + MVG_CT_XOFF + 7];
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
