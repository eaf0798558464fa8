"""Host front end (libmvfront.so): Annex-B / SPS / PPS / slice / CAVLC -> SoA."""
import numpy as np
import pytest

from helpers import golden_names, load_golden
from oracle import ref

FIELDS = ("mb_kind", "i16_mode", "chroma_mode", "qp_y", "cbp", "luma_modes", "coeff")


def same(a, b):
    return all(np.array_equal(getattr(a, f), getattr(b, f)) for f in FIELDS)


@pytest.mark.parametrize("name", golden_names())
def test_front_end_equals_the_reference_parse_on_golden_streams(name):
    """Golden fixtures hold the stream AND the reference's own parsed macroblocks + tables."""
    from minivideo_b200 import front
    want, z = load_golden(name)
    st = front.Stream(z["stream"].tobytes())
    assert (st.info.width_mbs, st.info.height_mbs, st.n_idr) == (want.width_mbs, want.height_mbs, want.n_pics)
    assert (st.info.cb_qp_offset, st.info.cr_qp_offset) == (want.cb_qp_offset, want.cr_qp_offset)
    ls4, ls8 = st.level_scale()
    assert np.array_equal(ls4, z["ls4"]) and np.array_equal(ls8, z["ls8"])
    for threads in (1, 3):
        assert same(st.parse(n_threads=threads), want)


@pytest.mark.parametrize("cfg,n", [("cif", 6), ("720p", 2), ("1080p", 2)])
def test_front_end_equals_generator_soa(cfg, n):
    from minivideo_b200 import front, synth
    stream, soa = synth.generate(n, cfg, seed=77)
    st = front.Stream(stream)
    assert same(st.parse(), soa)
    # sub-ranges and explicit index lists
    assert same(st.parse(first=1, count=n - 1), soa.pictures(1, n - 1))
    assert same(st.parse(indices=[n - 1, 0]), type(soa)(soa.width_mbs, soa.height_mbs, 2,
                *(np.concatenate([getattr(soa.pictures(n - 1, 1), f), getattr(soa.pictures(0, 1), f)]) for f in FIELDS)))


def test_three_byte_start_codes_and_other_nal_types_are_accepted():
    """The reference's ES parser needs 4-byte start codes and knows only 0x65/0x67/0x68
    (esparser.c:78-82); the front end takes the general Annex-B form."""
    from minivideo_b200 import front, synth
    stream, soa = synth.generate(2, "cif", seed=5)
    alt = stream.replace(b"\x00\x00\x00\x01", b"\x00\x00\x01")
    aud = b"\x00\x00\x01\x09\xf0"                       # access unit delimiter
    alt = aud + alt.replace(b"\x00\x00\x01\x65", aud + b"\x00\x00\x01\x25")   # nal_ref_idc 1
    assert same(front.Stream(alt).parse(), soa)


def test_errors_are_reported_not_fatal():
    from minivideo_b200 import front, synth
    with pytest.raises(front.FrontError):
        front.Stream(b"\x00" * 100)
    stream, _ = synth.generate(1, "cif")
    cut = stream[: len(stream) // 2]
    st = front.Stream(cut)
    with pytest.raises(front.FrontError, match="picture 0"):
        st.parse()
    with pytest.raises(front.FrontError, match="out of range"):
        front.Stream(stream).parse(indices=[3])
    cabac = bytearray(stream)
    i = stream.index(b"\x00\x00\x00\x01\x68") + 5
    cabac[i] |= 0x20                                     # ue(0) ue(0) then entropy_coding_mode_flag
    with pytest.raises(front.FrontError, match="CABAC") as e:
        front.Stream(bytes(cabac))
    assert e.value.code == -1


def _interleaved_stream(n, seed=9):
    """n small pictures whose coded sizes alternate between large and tiny."""
    from minivideo_b200 import synth
    kw = dict(width_mbs=4, height_mbs=3, profile_idc=66, seed=seed)
    big, _ = synth.generate(n, luma_cbp_percent=100, mean_coeffs_x10=120, **kw)
    small, _ = synth.generate(n, luma_cbp_percent=0, mean_coeffs_x10=1, **kw)
    def nals(s):
        parts = s[:-64].split(b"\x00\x00\x00\x01")[1:]
        return [b"\x00\x00\x00\x01" + p for p in parts]
    nb, ns = nals(big), nals(small)
    out = nb[0] + nb[1]
    for i in range(n):
        out += (ns if i % 3 == 1 else nb)[2 + i]
    return out + b"\x00" * 64


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("mode,flag", [(1, "ordered"), (2, "distributed")])
@pytest.mark.parametrize("n_total,n_want", [(20, 5), (60, 7), (31, 4)])
def test_idr_selection_matches_the_reference_filter(mode, flag, n_total, n_want, tmp_path):
    """mvf_select_idr() vs idr_filtering() (demuxer/filter.c): run the reference CLI with -e ordered /
    distributed and identify which pictures it exported."""
    import subprocess
    from minivideo_b200 import front
    stream = _interleaved_stream(n_total)
    st = front.Stream(stream)
    assert st.n_idr == n_total
    every = ref.decode(stream, n_total, 64, 48)["yuv"]
    (tmp_path / "in.264").write_bytes(stream)
    subprocess.run([str(ref.MINI_THUMBNAILER), "-i", "in.264", "-f", "yuv420", "-n", str(n_want), "-e", flag],
                   cwd=tmp_path, capture_output=True)
    got = []
    for i in range(n_want):
        f = tmp_path / f"in_{i}.yuv"
        if not f.exists():
            break
        pic = np.fromfile(f, np.uint8)
        hits = [k for k in range(n_total) if np.array_equal(every[k], pic)]
        assert hits, "reference exported a picture that is not in the stream"
        got.append(hits[0])
    mine = st.select_idr(n_want, mode).tolist()
    assert mine[:len(got)] == got and len(mine) >= len(got)
    assert len(got) >= n_want - 1            # the reference may lose the last one (index one past the end)


def test_unfiltered_selection_is_the_first_n():
    from minivideo_b200 import front, synth
    stream, _ = synth.generate(5, width_mbs=2, height_mbs=2, profile_idc=66)
    st = front.Stream(stream)
    assert st.select_idr(3, 0).tolist() == [0, 1, 2]
    assert st.select_idr(9, 0).tolist() == [0, 1, 2, 3, 4]
