"""Known-answer vectors the survey obtained from the reference:
uniform pictures whose value follows from one coefficient by hand arithmetic."""
import numpy as np
import pytest

from minivideo_b200.synth import Soa


def flat_soa(w, h, kind, qp, i16_mode=2, chroma_mode=0):
    n = w * h
    return Soa(w, h, 1, np.full(n, kind, np.uint8), np.full(n, i16_mode, np.uint8), np.full(n, chroma_mode, np.uint8),
               np.full(n, qp, np.int8), np.zeros(n, np.uint8), np.full((n, 16), 2, np.uint8), np.zeros((n, 384), np.int16))


@pytest.mark.parametrize("qp,expect", [(26, 130), (27, 131), (28, 131), (35, 135), (37, 136)])
def test_k1_k3_i16x16_dc_single_coefficient(qp, expect):
    """K1/K3: every MB Intra16x16 DC-predicted; MB0 carries one luma DC level +3.
    dcY = (3*LS + 2^(5-qP/6)) >> (6-qP/6), r = (dcY+32)>>6, Y = 128 + r everywhere."""
    from oracle import cpu
    soa = flat_soa(22, 18, 2, qp)
    soa.coeff[0, 0] = 3          # Intra16x16DCLevel[0] -> matrix (0,0) -> blk 0 slot 0
    yuv, _ = cpu.reconstruct(soa)
    y = yuv[0, :soa.width * soa.height]
    assert (y == expect).all(), np.unique(y)
    assert (yuv[0, soa.width * soa.height:] == 128).all()


def test_k2_i8x8_scaling_lists():
    """K2: High, 4x3 MBs, list i entry k = 16 + ((7i + 3k) mod 40); MB0 Intra8x8 with one level +3
    at the first coefficient of block 0: LS8 = 18*26 = 468, d = (3*468+2)>>2 = 351, r = 5 -> Y = 133."""
    from oracle import cpu
    soa = flat_soa(4, 3, 2, 26)
    soa.mb_kind[0] = 1
    soa.coeff[0, 0] = 3
    soa.lists4x4 = np.array([[16 + ((7 * i + 3 * k) % 40) for k in range(16)] for i in range(6)], np.uint8)
    soa.lists8x8 = np.array([[16 + ((7 * i + 3 * k) % 40) for k in range(64)] for i in (6, 7)], np.uint8)
    yuv, _ = cpu.reconstruct(soa)
    y = yuv[0, :soa.width * soa.height]
    assert (y == 133).all(), np.unique(y)


def test_rgb_formula_known_values():
    """export_utils.c:300-302 on hand-computed samples."""
    from oracle import cpu
    yuv = np.zeros((1, 16 * 16 * 3 // 2), np.uint8)
    yuv[0, :256] = 130
    yuv[0, 256:] = 128
    rgb = cpu.yuv_to_rgb(yuv, 16, 16, 1)
    t = (298 * 130) >> 8
    want = (t + ((408 * 128) >> 8) - 222, t - ((100 * 128) >> 8) - ((208 * 128) >> 8) + 135, t + ((516 * 128) >> 8) - 276)
    assert tuple(int(v) for v in rgb[0, 3, 5]) == want
    half = cpu.yuv_to_rgb(yuv, 16, 16, 2)
    assert half.shape == (1, 8, 8, 3) and tuple(int(v) for v in half[0, 0, 0]) == want
