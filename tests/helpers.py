"""Shared helpers for the tests."""
from __future__ import annotations

import hashlib
import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def golden_names():
    return sorted(p.stem for p in GOLDEN.glob("*.npz"))


def load_golden(name):
    """Returns (Soa built from the REFERENCE's parse, fixture dict)."""
    from minivideo_b200.synth import Soa
    z = np.load(GOLDEN / f"{name}.npz")
    soa = Soa(int(z["width_mbs"]), int(z["height_mbs"]), int(z["n_pics"]), z["soa_mb_kind"], z["soa_i16_mode"],
              z["soa_chroma_mode"], z["soa_qp_y"], z["soa_cbp"], z["soa_luma_modes"], z["soa_coeff"],
              cb_qp_offset=int(z["cb_qp_offset"]), cr_qp_offset=int(z["cr_qp_offset"]))
    return soa, z


def large_digests():
    return json.loads((GOLDEN / "large_digests.json").read_text())


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def oracle_sps_from_tables(soa, ls4, ls8):
    import ctypes as C
    from oracle import cpu
    s = cpu.OracleSps()
    s.width_mbs, s.height_mbs = soa.width_mbs, soa.height_mbs
    a = np.ascontiguousarray(ls4, np.int32); b = np.ascontiguousarray(ls8, np.int32)
    C.memmove(s.ls4, a.ctypes.data, a.nbytes)
    C.memmove(s.ls8, b.ctypes.data, b.nbytes)
    s.cb_qp_offset, s.cr_qp_offset = soa.cb_qp_offset, soa.cr_qp_offset
    return s


def oracle_reconstruct_with_tables(soa, ls4, ls8, want_residual=False):
    """Oracle reconstruction using explicit LevelScale tables (e.g. the reference's own)."""
    import ctypes as C
    from oracle import cpu
    L = cpu.lib()
    sps = oracle_sps_from_tables(soa, ls4, ls8)
    W, H, N = soa.width, soa.height, soa.n_mbs
    yuv = np.zeros((soa.n_pics, W * H * 3 // 2), np.uint8)
    res = np.zeros((soa.n_pics * N, 384), np.int16) if want_residual else None
    p_ = lambda a: a.ctypes.data_as(C.c_void_p)
    for p in range(soa.n_pics):
        s = slice(p * N, (p + 1) * N)
        arrs = [np.ascontiguousarray(x[s]) for x in (soa.mb_kind, soa.i16_mode, soa.chroma_mode, soa.qp_y, soa.luma_modes, soa.coeff)]
        y = yuv[p, :W * H]; cb = yuv[p, W * H:W * H * 5 // 4]; cr = yuv[p, W * H * 5 // 4:]
        r = res[s] if want_residual else None
        L.oracle_reconstruct_picture(C.byref(sps), *[p_(a) for a in arrs], p_(y), p_(cb), p_(cr),
                                     p_(r) if r is not None else None)
    return yuv, res
