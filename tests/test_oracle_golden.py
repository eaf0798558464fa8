"""The CPU oracle against the golden vectors produced by the unmodified reference
(tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from helpers import golden_names, large_digests, load_golden, oracle_reconstruct_with_tables, sha


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_output(name):
    from oracle import cpu
    soa, z = load_golden(name)
    yuv, _ = oracle_reconstruct_with_tables(soa, z["ls4"], z["ls8"])
    assert np.array_equal(yuv, z["yuv"]), f"{name}: oracle YUV differs from the reference's"
    rgb = cpu.yuv_to_rgb(yuv, soa.width, soa.height, 1)
    assert np.array_equal(rgb, z["rgb"]), f"{name}: oracle RGB differs from mb_to_rgb()"


@pytest.mark.parametrize("name", ["high_8x8_lists_qp0_51", "main_offsets", "cif_baseline"])
def test_level_scale_tables_match_reference(name):
    """oracle_build_level_scale / mvg_build_level_scale vs the reference's computeLevelScale*()."""
    from minivideo_b200 import api, synth
    from oracle import cpu
    soa, z = load_golden(name)
    # scaling lists are only known to the generator: re-create them from the fixture's stream params
    from golden.make_golden import SMALL
    n, kw = SMALL[name]
    _, gen = synth.generate(n, want_stream=False, **kw)
    for build in (cpu.level_scale, api.build_level_scale):
        ls4, ls8 = build(gen.lists4x4, gen.lists8x8[0])
        assert np.array_equal(ls4, z["ls4"]) and np.array_equal(ls8, z["ls8"])


@pytest.mark.parametrize("name", sorted(large_digests()))
def test_oracle_large_digests(name):
    """720p / 1080p: streams re-created from the seeded generator, outputs compared by SHA-256
    with what the reference produced when the fixtures were made."""
    from minivideo_b200 import synth
    from oracle import cpu
    d = large_digests()[name]
    stream, soa = synth.generate(d["n_pics"], **d["params"])
    assert sha(np.frombuffer(stream, np.uint8)) == d["stream_sha256"], "generator output drifted: re-run make_golden.py"
    yuv, _ = cpu.reconstruct(soa)
    rgb = cpu.yuv_to_rgb(yuv, soa.width, soa.height, 1)
    assert [sha(yuv[i]) for i in range(d["n_pics"])] == d["yuv_sha256"]
    assert [sha(rgb[i]) for i in range(d["n_pics"])] == d["rgb_sha256"]
