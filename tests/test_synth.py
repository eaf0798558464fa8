"""The synthetic CAVLC intra encoder (libmvsynth.so): determinism, constraints of SURVEY 8(c)."""
import numpy as np
import pytest


def test_deterministic_and_seed_sensitive():
    from minivideo_b200 import synth
    a, sa = synth.generate(2, "cif")
    b, sb = synth.generate(2, "cif")
    c, _ = synth.generate(2, "cif", seed=1)
    assert a == b and np.array_equal(sa.coeff, sb.coeff)
    assert a != c


def test_annexb_framing_obeys_the_reference_es_parser():
    """4-byte start codes, NAL header bytes 0x67/0x68/0x65 only, SPS+PPS first, 64 zero bytes of
    tail padding (esparser.c:65,:78-82), no start code emulation inside a NAL."""
    from minivideo_b200 import synth
    s, _ = synth.generate(3, "cif")
    assert s[:5] == b"\x00\x00\x00\x01\x67"
    assert s.endswith(b"\x00" * 64)
    body = s[:-64]
    starts = [i for i in range(len(body) - 4) if body[i:i + 4] == b"\x00\x00\x00\x01"]
    assert [body[i + 4] for i in starts] == [0x67, 0x68, 0x65, 0x65, 0x65]
    for i, st in enumerate(starts):          # emulation prevention: no 00 00 0x (x<=2) inside a NAL
        end = starts[i + 1] if i + 1 < len(starts) else len(body)
        nal = body[st + 4:end]
        for j in range(len(nal) - 2):
            assert not (nal[j] == 0 and nal[j + 1] == 0 and nal[j + 2] <= 2), (i, j)


def test_generator_avoids_the_reference_qp36_intra16x16_bug():
    """transform_16x16_lumadc() executes `1 << -1` at QP'Y == 36 (h264_transform.c:797-808)."""
    from minivideo_b200 import synth
    _, soa = synth.generate(4, want_stream=False, width_mbs=20, height_mbs=12, profile_idc=100, transform8x8=1,
                            qp_min=33, qp_max=39, init_qp=36, seed=5)
    assert (soa.qp_y == 36).any(), "test should exercise QP 36 on other MB kinds"
    assert not ((soa.qp_y == 36) & (soa.mb_kind == 2)).any()


def test_only_legal_prediction_modes_at_picture_borders():
    from minivideo_b200 import synth
    _, soa = synth.generate(2, want_stream=False, width_mbs=9, height_mbs=6, profile_idc=100, transform8x8=1, seed=6)
    W = soa.width_mbs
    for a in range(soa.n_mbs * 2):
        mx, my = (a % soa.n_mbs) % W, (a % soa.n_mbs) // W
        left, up = mx > 0, my > 0
        if soa.mb_kind[a] == 2:
            m = soa.i16_mode[a]
            assert not (m == 0 and not up) and not (m == 1 and not left) and not (m == 3 and not (left and up))
        cm = soa.chroma_mode[a]
        assert not (cm == 2 and not up) and not (cm == 1 and not left) and not (cm == 3 and not (left and up))
        if soa.mb_kind[a] == 0 and not left and not up:
            assert soa.luma_modes[a, 0] == 2       # only DC is legal for the very first block


def test_rejects_poc_types_the_reference_misparses():
    from minivideo_b200 import synth
    with pytest.raises(RuntimeError):
        synth.generate(1, "cif", poc_type=2)


def test_level_bounds_respected():
    from minivideo_b200 import synth
    _, soa = synth.generate(1, want_stream=False, width_mbs=10, height_mbs=8, profile_idc=77, max_level=9,
                            level_scale_x10=300, seed=12)
    ac = soa.coeff.copy()
    assert np.abs(ac).max() <= 4 * 9
