"""Regenerate the golden fixtures in this directory.

    python tests/golden/make_golden.py            (needs oracle/_ref, i.e. /root/reference mounted)

Every fixture is produced by the UNMODIFIED reference decoder (oracle/_ref/ref_decode,
built by oracle/Makefile from the reference sources): the Annex-B stream comes from the
committed synthetic encoder (minivideo_b200/csrc/h264_synth.c), the reference decodes it,
and we store
    stream      the Annex-B bytes (so the fixture does not depend on the generator)
    soa_*       the reference's own parsed macroblocks in the mvgpu.h SoA layout
    ls4, ls8    the reference's LevelScale tables, cb/cr offsets
    yuv, rgb    the reference's output planes / mb_to_rgb() output
Large cases (720p/1080p) only store SHA-256 digests of stream/yuv/rgb; their streams are
re-created from the seeded generator parameters recorded next to them.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from minivideo_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

SMALL = {
    "cif_baseline": (2, dict(config="cif")),
    "high_8x8_lists_qp0_51": (2, dict(width_mbs=9, height_mbs=7, profile_idc=100, transform8x8=1, scaling_lists=1,
                                      seed=3, qp_min=0, qp_max=51, init_qp=26, cb_qp_offset=4, cr_qp_offset=-5)),
    "main_offsets": (1, dict(width_mbs=8, height_mbs=6, profile_idc=77, seed=4, cb_qp_offset=-6, qp_min=12, qp_max=44)),
    "kind_i4x4": (1, dict(width_mbs=6, height_mbs=5, profile_idc=100, transform8x8=1, scaling_lists=1, force_kind=0, seed=100, qp_min=10, qp_max=45, init_qp=30)),
    "kind_i8x8": (1, dict(width_mbs=6, height_mbs=5, profile_idc=100, transform8x8=1, scaling_lists=0, force_kind=1, seed=101, qp_min=10, qp_max=45, init_qp=30)),
    "kind_i16x16": (1, dict(width_mbs=6, height_mbs=5, profile_idc=100, transform8x8=1, scaling_lists=1, force_kind=2, seed=102, qp_min=10, qp_max=45, init_qp=30)),
    "dense": (1, dict(width_mbs=8, height_mbs=8, profile_idc=100, transform8x8=1, scaling_lists=1, luma_cbp_percent=100,
                      mean_coeffs_x10=120, level_scale_x10=40, seed=7, qp_min=10, qp_max=40)),
    "big_levels": (1, dict(width_mbs=8, height_mbs=8, profile_idc=77, luma_cbp_percent=90, mean_coeffs_x10=60,
                           level_scale_x10=200, max_level=255, seed=8, qp_min=0, qp_max=51)),
    "pic_1x1": (3, dict(width_mbs=1, height_mbs=1, profile_idc=66, seed=9)),
    "pic_1xN": (2, dict(width_mbs=1, height_mbs=9, profile_idc=100, transform8x8=1, seed=10)),
    "pic_Nx1": (2, dict(width_mbs=13, height_mbs=1, profile_idc=100, transform8x8=1, seed=11)),
}
for m in range(9):
    SMALL[f"mode_{m}"] = (1, dict(width_mbs=5, height_mbs=4, profile_idc=100, transform8x8=1, force_mode=m, seed=200 + m))

LARGE = {
    "720p_main": (2, dict(config="720p")),
    "1080p_high": (2, dict(config="1080p")),
}


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    if not ref.available():
        raise SystemExit("oracle/_ref/ref_decode missing: run `make -C oracle ref` with the reference mounted")
    for name, (n, kw) in SMALL.items():
        stream, soa = synth.generate(n, **kw)
        r = ref.decode(stream, n, soa.width, soa.height, want_rgb=True, want_soa=True)
        rs, ls4, ls8 = ref.parse_soa(r["soa"])
        np.savez_compressed(
            HERE / f"{name}.npz", stream=np.frombuffer(stream, np.uint8), n_pics=n,
            width_mbs=rs.width_mbs, height_mbs=rs.height_mbs, cb_qp_offset=rs.cb_qp_offset, cr_qp_offset=rs.cr_qp_offset,
            ls4=ls4, ls8=ls8, soa_mb_kind=rs.mb_kind, soa_i16_mode=rs.i16_mode, soa_chroma_mode=rs.chroma_mode,
            soa_qp_y=rs.qp_y, soa_cbp=rs.cbp, soa_luma_modes=rs.luma_modes, soa_coeff=rs.coeff,
            yuv=r["yuv"], rgb=r["rgb"])
        print(f"{name:26s} {n} pics {soa.width}x{soa.height}  {(HERE / (name + '.npz')).stat().st_size / 1024:.0f} KiB")
    # a stream whose SPS / PPS change between pictures (tests/helpers.py builds it from three generated segments):
    # the reference re-decodes every parameter set where it meets it (h264.c:128-150)
    sys.path.insert(0, str(HERE.parent))
    from helpers import paramset_change_stream
    stream, segs = paramset_change_stream()
    n = sum(s.n_pics for s in segs)
    r = ref.decode(stream, n, segs[0].width, segs[0].height, want_rgb=True)
    np.savez_compressed(HERE / "multi_paramsets.npz", stream=np.frombuffer(stream, np.uint8), n_pics=n,
                        width_mbs=segs[0].width_mbs, height_mbs=segs[0].height_mbs, yuv=r["yuv"], rgb=r["rgb"])
    print(f"{'multi_paramsets':26s} {n} pics, 3 parameter generations")
    digests = {}
    for name, (n, kw) in LARGE.items():
        stream, soa = synth.generate(n, **kw)
        r = ref.decode(stream, n, soa.width, soa.height, want_rgb=True)
        digests[name] = dict(n_pics=n, params=kw, width=soa.width, height=soa.height,
                             stream_sha256=sha(np.frombuffer(stream, np.uint8)),
                             yuv_sha256=[sha(r["yuv"][i]) for i in range(n)],
                             rgb_sha256=[sha(r["rgb"][i]) for i in range(n)])
        print(f"{name:26s} {n} pics {soa.width}x{soa.height}  digests")
    (HERE / "large_digests.json").write_text(json.dumps(digests, indent=1))


if __name__ == "__main__":
    main()
