"""Randomised differential test: many small pictures with random geometry, profile, QP range and seed through the
CUDA path and through the oracle.  Small pictures put most macroblocks on a picture edge, where the availability
rules of every predictor (and kernel 2's table rows for one-sided DC, missing up-right neighbours and corner
replication) decide the result."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(20261018)
    for k in range(48):
        w, h = int(rng.integers(1, 7)), int(rng.integers(1, 7))
        high = bool(rng.integers(0, 2))
        kw = dict(width_mbs=w, height_mbs=h, seed=1000 + k, profile_idc=100 if high else 66,
                  transform8x8=int(high and rng.integers(0, 2)), scaling_lists=int(high and rng.integers(0, 2)),
                  qp_min=int(rng.integers(0, 20)), qp_max=int(rng.integers(30, 52)),
                  luma_cbp_percent=int(rng.choice([0, 50, 100])))
        if high:
            kw["cb_qp_offset"], kw["cr_qp_offset"] = int(rng.integers(-12, 13)), int(rng.integers(-12, 13))
        yield k, int(rng.integers(1, 9)), kw


@pytest.mark.parametrize("k,n,kw", list(_cases()), ids=[f"{k}-{kw['width_mbs']}x{kw['height_mbs']}" for k, _, kw in _cases()])
def test_random_small_pictures_match_the_oracle(k, n, kw):
    from minivideo_b200 import api, synth
    from oracle import cpu
    _, soa = synth.generate(n, want_stream=False, **kw)
    want_yuv, want_res = cpu.reconstruct(soa, want_residual=True)
    got = api.reconstruct(soa, rgb_scale=1, want_residual=True)
    assert np.array_equal(got["residual"], want_res)
    assert np.array_equal(got["yuv"], want_yuv)
    assert np.array_equal(got["rgb"], cpu.yuv_to_rgb(want_yuv, soa.width, soa.height, 1))
