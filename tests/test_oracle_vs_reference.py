"""Live check of the oracle (and of the synthetic encoder) against the compiled, unmodified
reference in oracle/_ref.  Skipped where oracle/_ref has not been built."""
import numpy as np
import pytest

from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (reference tree not mounted)")

SOA_FIELDS = ("mb_kind", "i16_mode", "chroma_mode", "qp_y", "cbp", "luma_modes", "coeff")

CASES = {
    "cif": (2, dict(config="cif", seed=31)),
    "high_small": (2, dict(width_mbs=7, height_mbs=5, profile_idc=100, transform8x8=1, scaling_lists=1, seed=32,
                           qp_min=0, qp_max=51, cb_qp_offset=-3, cr_qp_offset=5)),
    "main_dense": (1, dict(width_mbs=6, height_mbs=6, profile_idc=77, seed=33, luma_cbp_percent=100, mean_coeffs_x10=150)),
    "sparse": (1, dict(width_mbs=6, height_mbs=6, profile_idc=100, transform8x8=1, seed=34, luma_cbp_percent=5, mean_coeffs_x10=5)),
    "720p": (1, dict(config="720p", seed=35)),
    "1080p": (1, dict(config="1080p", seed=36)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_parse_equals_generator_soa_and_oracle_equals_reference(name):
    from minivideo_b200 import synth
    from oracle import cpu
    n, kw = CASES[name]
    stream, soa = synth.generate(n, **kw)
    r = ref.decode(stream, n, soa.width, soa.height, want_rgb=True, want_soa=True)
    parsed, ls4, ls8 = ref.parse_soa(r["soa"])
    for f in SOA_FIELDS:     # the reference parsed exactly the syntax the encoder meant
        assert np.array_equal(getattr(soa, f), getattr(parsed, f)), f
    o4, o8 = cpu.level_scale(soa.lists4x4, soa.lists8x8[0])
    assert np.array_equal(o4, ls4) and np.array_equal(o8, ls8)
    yuv, _ = cpu.reconstruct(soa)
    assert np.array_equal(yuv, r["yuv"])
    assert np.array_equal(cpu.yuv_to_rgb(yuv, soa.width, soa.height, 1), r["rgb"])


def test_export_tap_equals_the_reference_cli_files():
    """ref_decode's tap (mb_to_ycbcr) writes the same bytes as `mini_thumbnailer -f yuv420`
    (export_idr_yuv420 -> files in the CWD)."""
    from minivideo_b200 import synth
    stream, soa = synth.generate(3, "cif", seed=41)
    a = ref.decode(stream, 3, soa.width, soa.height)["yuv"]
    b = ref.decode_cli(stream, 3, soa.width, soa.height)
    assert np.array_equal(a, b)


def test_cavlc_coverage_every_total_coeff_and_nc_class():
    """Streams dense enough to reach every coeff_token table (nC classes 0-1, 2-3, 4-7, 8+ and
    chroma DC) and TotalCoeff 0..16 round-trip through the reference's CAVLC decoder."""
    from minivideo_b200 import synth
    seen = set()
    for seed, mean in ((51, 20), (52, 80), (53, 160)):
        stream, soa = synth.generate(1, width_mbs=8, height_mbs=6, profile_idc=77, seed=seed, luma_cbp_percent=90,
                                     mean_coeffs_x10=mean, force_kind=0)
        r = ref.decode(stream, 1, soa.width, soa.height, want_soa=True)
        parsed, _, _ = ref.parse_soa(r["soa"])
        assert np.array_equal(parsed.coeff, soa.coeff)
        seen |= set(np.count_nonzero(soa.coeff[:, :256].reshape(-1, 16), axis=1).tolist())
    assert seen == set(range(17))
