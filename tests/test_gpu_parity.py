"""Parity of the CUDA path (through the C ABI of libmvgpu.so) against the oracle and the golden
vectors of the reference.  Bit-exact: integer/byte work, no tolerance."""
import numpy as np
import pytest

from helpers import golden_names, large_digests, load_golden, sha

pytestmark = pytest.mark.gpu


def _ctx_for(soa, n_slots, ls4=None, ls8=None):
    from minivideo_b200 import api
    ctx = api.Context(0, soa.width_mbs, soa.height_mbs, n_slots)
    if ls4 is None:
        ctx.set_sps_from(soa)
    else:
        ctx.set_sps(soa.width_mbs, soa.height_mbs, ls4, ls8, soa.cb_qp_offset, soa.cr_qp_offset)
    return ctx


@pytest.mark.parametrize("name", golden_names())
def test_gpu_matches_reference_golden_vectors(name):
    """Input = the reference's own parsed macroblocks and LevelScale tables; expected output = the
    reference's own YUV and mb_to_rgb() bytes."""
    soa, z = load_golden(name)
    ctx = _ctx_for(soa, soa.n_pics, z["ls4"], z["ls8"])
    ctx.upload(soa, 0)
    ctx.run(0, soa.n_pics, 1)
    ctx.sync()
    for i in range(soa.n_pics):
        assert np.array_equal(ctx.download_yuv420(i), z["yuv"][i]), f"{name} pic {i} YUV"
        assert np.array_equal(ctx.download_rgb(i), z["rgb"][i]), f"{name} pic {i} RGB"
    ctx.run_rgb(0, soa.n_pics)                  # one fused kernel, levels -> RGB24
    ctx.sync()
    for i in range(soa.n_pics):
        assert np.array_equal(ctx.download_rgb(i), z["rgb"][i]), f"{name} pic {i} RGB (fused, direct)"
    ctx.close()


@pytest.mark.parametrize("name", sorted(large_digests()))
def test_gpu_large_digests(name):
    from minivideo_b200 import api, synth
    d = large_digests()[name]
    _, soa = synth.generate(d["n_pics"], want_stream=False, **d["params"])
    out = api.reconstruct(soa, rgb_scale=1)
    assert [sha(out["yuv"][i]) for i in range(d["n_pics"])] == d["yuv_sha256"]
    assert [sha(out["rgb"][i]) for i in range(d["n_pics"])] == d["rgb_sha256"]


CASES = {
    "cif_30": (30, dict(config="cif")),
    "720p": (3, dict(config="720p", seed=61)),
    "1080p": (4, dict(config="1080p", seed=62)),
    "qp_extremes": (2, dict(width_mbs=12, height_mbs=9, profile_idc=100, transform8x8=1, scaling_lists=1, seed=63, qp_min=0, qp_max=51)),
    "big_levels": (2, dict(width_mbs=12, height_mbs=9, profile_idc=100, transform8x8=1, seed=64, level_scale_x10=300, max_level=255, qp_min=0, qp_max=51, luma_cbp_percent=95)),
    "hostile_levels": (1, dict(width_mbs=8, height_mbs=8, profile_idc=100, transform8x8=1, scaling_lists=1, seed=65, level_scale_x10=30000, max_level=8191, qp_min=40, qp_max=51, init_qp=45, luma_cbp_percent=100)),
    "one_mb": (5, dict(width_mbs=1, height_mbs=1, profile_idc=100, transform8x8=1, seed=66)),
    "one_row": (2, dict(width_mbs=33, height_mbs=1, profile_idc=100, transform8x8=1, seed=67)),
    "one_col": (2, dict(width_mbs=1, height_mbs=33, profile_idc=100, transform8x8=1, seed=68)),
    "two_cols": (2, dict(width_mbs=2, height_mbs=17, profile_idc=100, transform8x8=1, seed=69)),
    "2160p": (1, dict(config="2160p", seed=70)),
}


@pytest.mark.parametrize("mode", ["fused", "split"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_oracle(name, mode):
    """Same seeded SoA through the oracle and the CUDA path: residual (the transform stage), YUV and RGB
    must be identical, including int32 wrap-around on hostile levels.  `fused`: kf_recon (tiles, then
    kernel 3) and kf_recon straight to RGB24; `split`: round 1's kernels 1, 2, 3."""
    from minivideo_b200 import api, synth
    from oracle import cpu
    n, kw = CASES[name]
    _, soa = synth.generate(n, want_stream=False, **kw)
    want_yuv, want_res = cpu.reconstruct(soa, want_residual=True)
    got = api.reconstruct(soa, rgb_scale=1, want_residual=True,
                          mode=api.PIPELINE_FUSED if mode == "fused" else api.PIPELINE_SPLIT)
    assert np.array_equal(got["residual"], want_res)
    assert np.array_equal(got["yuv"], want_yuv)
    want_rgb = cpu.yuv_to_rgb(want_yuv, soa.width, soa.height, 1)
    assert np.array_equal(got["rgb_k3"], want_rgb)
    assert np.array_equal(got["rgb"], want_rgb)
    assert got["timing"].launches == (2 if mode == "fused" else 3)


@pytest.mark.parametrize("geometry", [(12, 8), (13, 5), (1, 3)])
@pytest.mark.parametrize("scale", [2, 4, 8, 16])
def test_gpu_rgb_downscale_matches_oracle_box_filter(scale, geometry):
    """Thumbnails through both routes -- mvg_run_thumbs() (the fused kernel's thumbnail mode: levels -> RGB24 at 1/scale)
    and mvg_run() (tiles, then kernel 3) -- against the oracle's box filter; widths that are not a multiple of the
    kernel's group of four macroblocks included.  And the end-to-end call, which takes the one-kernel route."""
    from minivideo_b200 import api, synth
    from oracle import cpu
    w, h = geometry
    _, soa = synth.generate(2, want_stream=False, width_mbs=w, height_mbs=h, profile_idc=100, transform8x8=1, seed=71 + w)
    want_yuv, _ = cpu.reconstruct(soa)
    want = cpu.yuv_to_rgb(want_yuv, soa.width, soa.height, scale)
    got = api.reconstruct(soa, rgb_scale=scale)
    assert np.array_equal(got["rgb"], want)
    assert np.array_equal(got["rgb_k3"], want)
    ctx = _ctx_for(soa, soa.n_pics)
    try:
        out = np.zeros_like(want)
        ctx.decode_host(soa, None, out, rgb_scale=scale)
        assert np.array_equal(out, want)
    finally:
        ctx.close()


@pytest.mark.parametrize("scale", [3, 6, 12])
def test_gpu_rgb_downscale_by_other_factors_uses_the_generic_kernel(scale):
    """Scales that are not a power of two (or exceed 16) go through k3_rgb_scaled_generic."""
    from minivideo_b200 import api, synth
    from oracle import cpu
    _, soa = synth.generate(2, want_stream=False, width_mbs=6, height_mbs=3, profile_idc=100, transform8x8=1, seed=72)
    want_yuv, _ = cpu.reconstruct(soa)
    got = api.reconstruct(soa, rgb_scale=scale)
    assert np.array_equal(got["rgb"], cpu.yuv_to_rgb(want_yuv, soa.width, soa.height, scale))


def test_full_size_batch_properties_1080p():
    """BASELINE-size property checks (the oracle would take minutes on this many pictures):
    a batch built from 4 distinct pictures cloned into 96 slots reconstructs every clone to the
    same bytes as its source, whatever wave of the persistent kernel processed it; a second run is
    bit-identical (no race); RGB at scale 4 equals the box average of the scale-1 RGB."""
    from minivideo_b200 import api, synth
    from oracle import cpu
    G, F = 4, 96
    _, soa = synth.generate(G, want_stream=False, config="1080p", seed=81)
    ctx = _ctx_for(soa, F)
    ctx.upload(soa, 0)
    for s in range(G, F):
        ctx.clone_slot(s % G, s)
    ctx.run(0, F, 1)
    ctx.sync()
    want_yuv, _ = cpu.reconstruct(soa.pictures(0, 2))
    base = [ctx.download_yuv420(i) for i in range(G)]
    base_rgb = [ctx.download_rgb(i) for i in range(G)]
    assert np.array_equal(base[0], want_yuv[0]) and np.array_equal(base[1], want_yuv[1])
    for s in (G, G + 1, 37, 50, F - 2, F - 1):
        assert np.array_equal(ctx.download_yuv420(s), base[s % G]), s
        assert np.array_equal(ctx.download_rgb(s), base_rgb[s % G]), s
    ctx.run(0, F, 4)
    ctx.sync()
    for s in (0, 1, F - 1):
        assert np.array_equal(ctx.download_yuv420(s), base[s % G])
        full = base_rgb[s % G].astype(np.uint32)
        h, w = full.shape[:2]
        box = (full.reshape(h // 4, 4, w // 4, 4, 3).sum(axis=(1, 3)) + 8) // 16
        assert np.array_equal(ctx.download_rgb(s, 4), box.astype(np.uint8))
    t = ctx.timing()
    assert t.launches == 2 and t.fused_ms > 0 and t.k3_rgb_ms > 0       # kf_recon<tiles>, k3_rgb_scaled
    # the one-kernel path (levels -> RGB24) over the same batch: every clone equals its source, and a
    # repeat is bit-identical
    ctx.run_rgb(0, F)
    ctx.sync()
    assert ctx.timing().launches == 1
    for s in (0, 1, G, 37, 50, F - 2, F - 1):
        assert np.array_equal(ctx.download_rgb(s), base_rgb[s % G]), s
    with pytest.raises(api.MvgError, match="RGB24 only"):
        ctx.download_yuv420(0)
    ctx.close()


def test_decode_host_equals_resident_path_and_handles_ragged_batches():
    """mvg_decode_host (pinned H2D/D2H pipeline over slot regions) with batch sizes that do not
    divide the region size, including a batch larger than the context."""
    from minivideo_b200 import api, synth
    from oracle import cpu
    _, soa = synth.generate(11, want_stream=False, width_mbs=10, height_mbs=7, profile_idc=100, transform8x8=1,
                            scaling_lists=1, seed=91)
    want_yuv, _ = cpu.reconstruct(soa)
    want_rgb = cpu.yuv_to_rgb(want_yuv, soa.width, soa.height, 1)
    for slots in (1, 2, 4, 6, 16):
        ctx = _ctx_for(soa, slots)
        yuv = np.zeros_like(want_yuv)
        rgb = np.zeros((soa.n_pics, soa.height * soa.width * 3), np.uint8)
        ctx.decode_host(soa, yuv, rgb, 1)
        assert np.array_equal(yuv, want_yuv), slots
        assert np.array_equal(rgb.reshape(want_rgb.shape), want_rgb), slots
        ctx.close()


def test_error_paths():
    from minivideo_b200 import api, synth
    _, soa = synth.generate(1, want_stream=False, width_mbs=4, height_mbs=4, profile_idc=66)
    ctx = api.Context(0, 4, 4, 2)
    with pytest.raises(api.MvgError, match="mvg_set_sps"):
        ctx.run(0, 1, 1)
    ctx.set_sps_from(soa)
    with pytest.raises(api.MvgError, match="slots"):
        ctx.run(1, 2, 1)
    with pytest.raises(api.MvgError, match="rgb_scale"):
        ctx.upload(soa, 0); ctx.run(0, 1, 3)
    with pytest.raises(api.MvgError, match="capacity"):
        ctx.set_sps(5, 4, *api.build_level_scale(None, None))
    ctx.close()
    with pytest.raises(api.MvgError, match="out of range"):
        api.Context(99, 4, 4, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [9, 10, 11, 12, 13, 14])
def test_garbage_side_information_terminates_and_leaves_the_context_usable(seed):
    """Random bytes in every SoA field (macroblock kinds, modes and QPs far outside their ranges, arbitrary levels): the
    result is undefined, but the kernels must finish (every macroblock still publishes its line, so no row waits for
    ever), must not fault, and the context must reconstruct a proper batch correctly afterwards."""
    from minivideo_b200 import api, synth
    from oracle import cpu
    _, good = synth.generate(3, want_stream=False, width_mbs=9, height_mbs=7, profile_idc=100, transform8x8=1, seed=71)
    rng = np.random.default_rng(seed)
    bad = synth.generate(3, want_stream=False, width_mbs=9, height_mbs=7, profile_idc=100, transform8x8=1, seed=72)[1]
    for name in ("mb_kind", "i16_mode", "chroma_mode", "cbp", "luma_modes"):
        a = getattr(bad, name)
        a[...] = rng.integers(0, 256, a.shape, dtype=np.uint8)
    bad.qp_y[...] = rng.integers(-128, 128, bad.qp_y.shape).astype(np.int8)
    bad.coeff[...] = rng.integers(-32768, 32768, bad.coeff.shape).astype(np.int16)
    ctx = api.Context(0, 9, 7, 3)
    ctx.set_sps_from(good)
    ctx.upload(bad, 0)
    ctx.run(0, 3, 1)
    ctx.sync()                                          # a fault or a hang would surface here
    ctx.upload(good, 0)
    ctx.run(0, 3, 1)
    ctx.sync()
    want, _ = cpu.reconstruct(good, want_residual=True)
    for i in range(3):
        assert np.array_equal(ctx.download_yuv420(i), want[i])
    ctx.close()


def test_run_thumbs_contract():
    """mvg_run_thumbs(): scale 1 is the full-size one-kernel path, a divisor that is not 2, 4, 8, 16 falls back to tiles +
    kernel 3 (planar picture available afterwards), the thumbnail mode leaves no planar picture behind, and a scale that
    does not divide the picture is refused with a message.  Random garbage through the thumbnail mode terminates."""
    from minivideo_b200 import api, synth
    from oracle import cpu
    _, soa = synth.generate(2, want_stream=False, width_mbs=9, height_mbs=6, profile_idc=100, transform8x8=1, seed=91)
    want_yuv, _ = cpu.reconstruct(soa)
    ctx = _ctx_for(soa, soa.n_pics)
    try:
        ctx.upload(soa, 0)
        for scale in (1, 2, 3, 6, 8):                                       # 144 x 96: 3 and 6 divide it, 8 too
            ctx.run_thumbs(0, soa.n_pics, scale)
            ctx.sync()
            want = cpu.yuv_to_rgb(want_yuv, soa.width, soa.height, scale)
            got = np.stack([ctx.download_rgb(i, scale) for i in range(soa.n_pics)])
            assert np.array_equal(got, want), scale
            if scale in (3, 6):
                assert np.array_equal(ctx.download_yuv420(0), want_yuv[0])
            else:
                with pytest.raises(api.MvgError, match="RGB24 only"):
                    ctx.download_yuv420(0)
        with pytest.raises(api.MvgError, match="rgb_scale"):
            ctx.run_thumbs(0, soa.n_pics, 5)
        with pytest.raises(api.MvgError, match="rgb_scale"):
            ctx.run_thumbs(0, soa.n_pics, 0)
        rng = np.random.default_rng(5)
        bad = synth.generate(2, want_stream=False, width_mbs=9, height_mbs=6, profile_idc=100, transform8x8=1, seed=92)[1]
        for name in ("mb_kind", "i16_mode", "chroma_mode", "cbp", "luma_modes"):
            a = getattr(bad, name)
            a[...] = rng.integers(0, 256, a.shape, dtype=np.uint8)
        bad.qp_y[...] = rng.integers(-128, 128, bad.qp_y.shape).astype(np.int8)
        ctx.upload(bad, 0)
        ctx.run_thumbs(0, bad.n_pics, 4)
        ctx.sync()                                                          # a fault or a hang would surface here
        ctx.upload(soa, 0)
        ctx.run_thumbs(0, soa.n_pics, 4)
        ctx.sync()
        assert np.array_equal(ctx.download_rgb(1, 4), cpu.yuv_to_rgb(want_yuv, soa.width, soa.height, 4)[1])
    finally:
        ctx.close()


def test_async_submissions_two_batches_in_flight():
    """mvg_submit*/mvg_wait (the asynchronous boundary): several submissions in flight through one context -- different
    batches, packed and dense, RGB and YUV, a context smaller than a batch -- complete with the right bytes whatever the
    order they are waited for in; mvg_poll answers without blocking; a ticket cannot be waited for twice."""
    from minivideo_b200 import api, synth
    from oracle import cpu
    kw = dict(width_mbs=20, height_mbs=12, profile_idc=100, transform8x8=1, scaling_lists=1)
    _, a = synth.generate(7, want_stream=False, seed=91, **kw)
    _, b = synth.generate(5, want_stream=False, seed=91, **kw)          # same tables (lists come from the seed)
    b = b.pictures(2, 3)
    want = {"a": cpu.reconstruct(a)[0], "b": cpu.reconstruct(b)[0]}
    want_rgb = {k: cpu.yuv_to_rgb(v, a.width, a.height, 1) for k, v in want.items()}
    for slots in (4, 16):
        ctx = _ctx_for(a, slots)
        pa, pb = api.Packed(a, pinned=True), api.Packed(b, pinned=True)
        rgb_a = api.PinnedArray((a.n_pics, a.height, a.width, 3), np.uint8)
        yuv_b = api.PinnedArray((b.n_pics, a.width * a.height * 3 // 2), np.uint8)
        rgb_b = api.PinnedArray((b.n_pics, a.height, a.width, 3), np.uint8)
        rgb_a2 = api.PinnedArray((a.n_pics, a.height, a.width, 3), np.uint8)
        t1 = api.submit_packed(ctx, pa, None, rgb_a.array, 1)
        t2 = api.submit_packed(ctx, pb, yuv_b.array, rgb_b.array, 1)
        keep = []
        t3 = ctx.submit(a, None, rgb_a2.array, 1, keep)                 # dense levels, third submission in flight
        assert len({t1, t2, t3}) == 3
        ctx.wait(t2)                                                    # out of order
        assert np.array_equal(yuv_b.array, want["b"]) and np.array_equal(rgb_b.array, want_rgb["b"])
        ctx.wait(t3)
        assert ctx.poll(t1) is True                                     # submitted before t3: done by now
        ctx.wait(t1)
        assert np.array_equal(rgb_a.array, want_rgb["a"]) and np.array_equal(rgb_a2.array, want_rgb["a"])
        with pytest.raises(api.MvgError, match="not in flight"):
            ctx.wait(t1)
        # the blocking calls are submit + wait: same results right after
        rgb_a.array[...] = 0
        ctx.decode_host_packed(pa, None, rgb_a.array.reshape(a.n_pics, -1), 1)
        assert np.array_equal(rgb_a.array, want_rgb["a"])
        # more than eight submissions in flight are refused with a message, and the context stays usable
        tickets = [api.submit_packed(ctx, pb, None, rgb_b.array, 1) for _ in range(8)]
        with pytest.raises(api.MvgError, match="too many submissions"):
            api.submit_packed(ctx, pb, None, rgb_b.array, 1)
        for t in tickets:
            ctx.wait(t)
        assert np.array_equal(rgb_b.array, want_rgb["b"])
        ctx.close()
