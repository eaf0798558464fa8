"""compute-sanitizer workload: the smoke() invocation plus the small geometries that exercise every edge of the
wavefront kernel (one row, one column, one macroblock, CIF, a small High-profile picture with 8x8 blocks and
scaling lists), each checked against the oracle.  Run under
    compute-sanitizer --tool {memcheck,racecheck,synccheck,initcheck} python tests/tools/sanitize_cases.py
(scripts/sanitize.sh does all four and keeps the logs)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from minivideo_b200 import api, synth  # noqa: E402
from oracle import cpu  # noqa: E402

CASES = [
    (2, dict(config="cif")),
    (1, dict(width_mbs=10, height_mbs=6, profile_idc=100, transform8x8=1, scaling_lists=1, cb_qp_offset=2, cr_qp_offset=-3, seed=77)),
    (2, dict(width_mbs=1, height_mbs=1, profile_idc=100, transform8x8=1, seed=66)),
    (1, dict(width_mbs=33, height_mbs=1, profile_idc=100, transform8x8=1, seed=67)),
    (1, dict(width_mbs=1, height_mbs=33, profile_idc=100, transform8x8=1, seed=68)),
    (1, dict(width_mbs=2, height_mbs=17, profile_idc=100, transform8x8=1, seed=69)),
]
for n, kw in CASES:
    _, soa = synth.generate(n, want_stream=False, **kw)
    want = cpu.reconstruct(soa)[0]
    for scale in (1, 4) if soa.width % 4 == 0 and soa.height % 4 == 0 else (1,):
        got = api.reconstruct(soa, device=0, rgb_scale=scale)
        assert np.array_equal(got["yuv"], want), kw
        assert np.array_equal(got["rgb"], cpu.yuv_to_rgb(want, soa.width, soa.height, scale)), kw
    # the end-to-end entry points (packed levels: kernel 0 too)
    ctx = api.Context(0, soa.width_mbs, soa.height_mbs, max(2, n))
    ctx.set_sps_from(soa)
    yuv = np.zeros_like(want)
    rgb = np.zeros((n, soa.width * soa.height * 3), np.uint8)
    ctx.decode_host_packed(api.Packed(soa), yuv, rgb, 1)
    assert np.array_equal(yuv, want), kw
    assert np.array_equal(rgb, cpu.yuv_to_rgb(want, soa.width, soa.height, 1)), kw
    ctx.close()
print("sanitize_cases: all cases bit-exact")
