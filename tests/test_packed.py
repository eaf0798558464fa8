"""Packed transfer format (mvg_packed_batch, include/mvgpu.h): host packer against a numpy restatement,
and the packed end-to-end entry point against the dense one and the oracle."""
import ctypes as C

import numpy as np
import pytest


def _soa(n=3, **kw):
    from minivideo_b200 import synth
    args = dict(width_mbs=7, height_mbs=5, profile_idc=100, transform8x8=1, scaling_lists=1, seed=311)
    args.update(kw)
    return synth.generate(n, want_stream=False, **args)[1]


def test_pack_round_trip_and_layout():
    from minivideo_b200 import api
    soa = _soa()
    pk = api.Packed(soa)
    assert np.array_equal(api.unpack_levels(pk), soa.coeff)
    # bitmap = chunks with a non-zero level; offsets are running sums inside a picture
    nz = (soa.coeff.reshape(-1, 24, 16) != 0)
    want_nzb = (nz.any(axis=2) * (1 << np.arange(24))).sum(axis=1).astype(np.uint32)
    assert np.array_equal(pk.nz_blocks, want_nzb)
    per_mb = nz.any(axis=2).sum(axis=1) + nz.sum(axis=(1, 2))
    per_pic = per_mb.reshape(soa.n_pics, -1)
    assert np.array_equal(pk.word_off.reshape(soa.n_pics, -1), np.cumsum(per_pic, axis=1) - per_pic)
    assert np.array_equal(pk.pic_off, np.concatenate([[0], np.cumsum(per_pic.sum(axis=1))]).astype(np.uint64))
    assert pk.nbytes < soa.coeff.nbytes / 3


@pytest.mark.parametrize("threads", [1, 2, 5])
def test_pack_is_thread_count_independent_and_handles_extremes(threads):
    from minivideo_b200 import api
    soa = _soa(4)
    soa.coeff[: soa.n_mbs] = 0                                    # an empty picture
    soa.coeff[soa.n_mbs: 2 * soa.n_mbs] = np.arange(1, 385, dtype=np.int16) - 200   # (almost) every level set
    soa.coeff[soa.n_mbs, 200] = -32768
    pk = api.Packed(soa, n_threads=threads)
    assert np.array_equal(api.unpack_levels(pk), soa.coeff)
    assert int(pk.pic_off[1]) == 0
    assert int(pk.pic_off[2] - pk.pic_off[1]) == soa.n_mbs * api.WORDS_PER_MB - soa.n_mbs   # one zero level per MB (index 199)


def test_pack_reports_a_short_buffer():
    from minivideo_b200 import api
    lib = api.load_library()
    soa = _soa(1)
    n = soa.n_mbs
    nzb, off, po, words = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(2, np.uint64), np.zeros(8, np.uint16)
    rc = lib.mvg_pack_batch(soa.coeff.ctypes.data, 1, n, nzb.ctypes.data, off.ctypes.data, po.ctypes.data,
                            words.ctypes.data, words.size, 1)
    assert rc == 0


@pytest.mark.gpu
@pytest.mark.parametrize("slots", [1, 3, 16])
def test_packed_decode_equals_dense_decode_and_oracle(slots):
    from minivideo_b200 import api
    from oracle import cpu
    soa = _soa(11, cb_qp_offset=2, cr_qp_offset=-3)
    soa.coeff[3 * soa.n_mbs: 4 * soa.n_mbs] = 0                   # a picture without residual
    want_yuv, _ = cpu.reconstruct(soa)
    want_rgb = cpu.yuv_to_rgb(want_yuv, soa.width, soa.height, 1)
    ctx = api.Context(0, soa.width_mbs, soa.height_mbs, slots)
    ctx.set_sps_from(soa)
    pk = api.Packed(soa, pinned=True)
    yuv = np.zeros_like(want_yuv)
    rgb = np.zeros((soa.n_pics, soa.height * soa.width * 3), np.uint8)
    ctx.decode_host_packed(pk, yuv, rgb, 1)
    assert np.array_equal(yuv, want_yuv)
    assert np.array_equal(rgb.reshape(want_rgb.shape), want_rgb)
    yuv2 = np.zeros_like(want_yuv)
    ctx.decode_host(soa, yuv2, None, 0)
    assert np.array_equal(yuv2, yuv)
    # thumbnails through the packed path
    small = np.zeros((soa.n_pics, (soa.height // 4) * (soa.width // 4) * 3), np.uint8)
    ctx.decode_host_packed(pk, None, small, 4)
    assert np.array_equal(small.reshape(soa.n_pics, soa.height // 4, soa.width // 4, 3),
                          cpu.yuv_to_rgb(want_yuv, soa.width, soa.height, 4))
    ctx.close()


@pytest.mark.gpu
def test_packed_decode_rejects_bad_batches():
    from minivideo_b200 import api
    soa = _soa(2)
    ctx = api.Context(0, soa.width_mbs, soa.height_mbs, 2)
    pk = api.Packed(soa)
    out = np.zeros((2, soa.height * soa.width * 3), np.uint8)
    with pytest.raises(api.MvgError, match="mvg_set_sps"):
        ctx.decode_host_packed(pk, None, out, 1)
    ctx.set_sps_from(soa)
    pk.pic_off[1] = pk.pic_off[2] + 5
    with pytest.raises(api.MvgError, match="impossible word count"):
        ctx.decode_host_packed(pk, None, out, 1)
    ctx.close()


@pytest.mark.parametrize("threads", [1, 4])
def test_front_end_emits_the_same_packed_batch_as_the_packer(threads):
    """mvf_parse_pictures_packed(stream) == mvg_pack_batch(mvf_parse_pictures(stream)), array by array."""
    from minivideo_b200 import api, front, synth
    stream, soa = synth.generate(5, width_mbs=9, height_mbs=6, profile_idc=100, transform8x8=1, scaling_lists=1, seed=321)
    st = front.Stream(stream)
    got = st.parse_packed(n_threads=threads)
    want = api.Packed(st.parse())
    for k in ("mb_kind", "i16_mode", "chroma_mode", "qp_y", "luma_modes", "nz_blocks", "word_off", "pic_off"):
        assert np.array_equal(got[k], getattr(want, k)), k
    n = int(want.pic_off[-1])
    assert np.array_equal(got["words"][:n], want.words[:n])
    # sub-range with explicit indices, and the capacity check
    sub = st.parse_packed(indices=[3, 0], n_threads=threads)
    w2 = api.Packed(st.parse(indices=[3, 0]))
    assert np.array_equal(sub["words"][: int(w2.pic_off[-1])], w2.words[: int(w2.pic_off[-1])])
    with pytest.raises(front.FrontError, match="capacity"):
        st.parse_packed(words_capacity=10)


@pytest.mark.gpu
def test_packed_decode_survives_malformed_offsets_and_masks():
    """Offsets and masks that point past a picture's words must not fault the device: the expand kernel
    bounds every read by the picture's word range (results are then unspecified, not checked)."""
    from minivideo_b200 import api
    soa = _soa(3)
    ctx = api.Context(0, soa.width_mbs, soa.height_mbs, 3)
    ctx.set_sps_from(soa)
    pk = api.Packed(soa, pinned=True)
    out = np.zeros((3, soa.height * soa.width * 3), np.uint8)
    pk.word_off[5] = 0xFFFFFF00                       # far outside
    pk.word_off[soa.n_mbs + 1] = int(pk.pic_off[2] - pk.pic_off[1])      # exactly at the end of picture 1
    pk.nz_blocks[7] = 0x00FFFFFF                      # claims 24 coded chunks
    pk.words[: 64] = 0xFFFF                           # masks claiming 16 levels each
    ctx.decode_host_packed(pk, None, out, 1)
    good = api.Packed(soa, pinned=True)               # the context is still usable afterwards
    ctx.decode_host_packed(good, None, out, 1)
    from oracle import cpu
    want = cpu.yuv_to_rgb(cpu.reconstruct(soa)[0], soa.width, soa.height, 1)
    assert np.array_equal(out.reshape(want.shape), want)
    ctx.close()
