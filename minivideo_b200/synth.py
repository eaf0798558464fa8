"""ctypes binding of libmvsynth.so: the committed minimal CAVLC intra encoder that
produces the synthetic Annex-B streams (and the matching SoA) used by the tests
and the benchmark.  See include/mvsynth.h."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

_LIB = None


class _Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "width_mbs", "height_mbs", "n_pics", "profile_idc", "transform8x8", "scaling_lists",
        "w_i4x4", "w_i8x8", "w_i16x16", "init_qp", "qp_min", "qp_max",
        "cb_qp_offset", "cr_qp_offset", "luma_cbp_percent", "mean_coeffs_x10",
        "level_scale_x10", "max_level", "poc_type", "crop_bottom", "force_mode", "force_kind")]
    _fields_.append(("seed", C.c_uint64))


class _Output(C.Structure):
    _fields_ = [
        ("stream", C.c_void_p), ("stream_cap", C.c_size_t), ("stream_len", C.c_size_t),
        ("mb_kind", C.c_void_p), ("i16_mode", C.c_void_p), ("chroma_mode", C.c_void_p),
        ("cbp", C.c_void_p), ("luma_modes", C.c_void_p), ("qp_y", C.c_void_p), ("coeff", C.c_void_p),
        ("lists4x4", (C.c_uint8 * 16) * 6), ("lists8x8", (C.c_uint8 * 64) * 2),
    ]


def _lib():
    global _LIB
    if _LIB is None:
        path = Path(__file__).resolve().parent / "libmvsynth.so"
        if not path.exists():
            from . import build
            build.build_synth()
        _LIB = C.CDLL(str(path))
        _LIB.mvs_default_params.argtypes = [C.POINTER(_Params)]
        _LIB.mvs_generate.argtypes = [C.POINTER(_Params), C.POINTER(_Output)]
        _LIB.mvs_generate.restype = C.c_int
        _LIB.mvs_stream_bound.argtypes = [C.POINTER(_Params)]
        _LIB.mvs_stream_bound.restype = C.c_size_t
    return _LIB


@dataclass
class Soa:
    """One batch of parsed pictures in the mvgpu.h structure-of-arrays layout."""
    width_mbs: int
    height_mbs: int
    n_pics: int
    mb_kind: np.ndarray      # u8  [P*N]
    i16_mode: np.ndarray     # u8  [P*N]
    chroma_mode: np.ndarray  # u8  [P*N]
    qp_y: np.ndarray         # i8  [P*N]
    cbp: np.ndarray          # u8  [P*N]
    luma_modes: np.ndarray   # u8  [P*N,16]
    coeff: np.ndarray        # i16 [P*N,384]
    lists4x4: np.ndarray = field(default_factory=lambda: np.full((6, 16), 16, np.uint8))
    lists8x8: np.ndarray = field(default_factory=lambda: np.full((2, 64), 16, np.uint8))
    cb_qp_offset: int = 0
    cr_qp_offset: int = 0

    @property
    def n_mbs(self) -> int:
        return self.width_mbs * self.height_mbs

    @property
    def width(self) -> int:
        return 16 * self.width_mbs

    @property
    def height(self) -> int:
        return 16 * self.height_mbs

    def pictures(self, first: int, count: int) -> "Soa":
        n = self.n_mbs
        sl = slice(first * n, (first + count) * n)
        return Soa(self.width_mbs, self.height_mbs, count, self.mb_kind[sl], self.i16_mode[sl],
                   self.chroma_mode[sl], self.qp_y[sl], self.cbp[sl], self.luma_modes[sl], self.coeff[sl],
                   self.lists4x4, self.lists8x8, self.cb_qp_offset, self.cr_qp_offset)


# configs of BASELINE.json / SURVEY.md section 8(d); seed = 0xC0FFEE + config index
CONFIGS = {
    "cif":   dict(width_mbs=22, height_mbs=18, profile_idc=66, seed=0xC0FFEE + 0),
    "720p":  dict(width_mbs=80, height_mbs=45, profile_idc=77, seed=0xC0FFEE + 1),
    "1080p": dict(width_mbs=120, height_mbs=68, profile_idc=100, transform8x8=1, scaling_lists=1,
                  crop_bottom=4, cb_qp_offset=2, cr_qp_offset=-2, seed=0xC0FFEE + 2),
    "2160p": dict(width_mbs=240, height_mbs=135, profile_idc=100, transform8x8=1, scaling_lists=1,
                  cb_qp_offset=2, cr_qp_offset=-2, seed=0xC0FFEE + 4),
}


def default_params(**overrides) -> _Params:
    p = _Params()
    _lib().mvs_default_params(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def generate(n_pics: int = 1, config: str | None = None, want_stream: bool = True,
             want_soa: bool = True, **overrides) -> tuple[bytes | None, Soa | None]:
    """Generate `n_pics` intra pictures.  Returns (annexb_bytes, Soa)."""
    kw = dict(CONFIGS[config]) if config else {}
    kw.update(overrides)
    kw["n_pics"] = n_pics
    p = default_params(**kw)
    lib = _lib()
    out = _Output()
    n = p.width_mbs * p.height_mbs * n_pics
    stream = None
    if want_stream:
        cap = lib.mvs_stream_bound(C.byref(p))
        stream = np.empty(cap, np.uint8)
        out.stream = stream.ctypes.data
        out.stream_cap = cap
    soa = None
    if want_soa:
        soa = Soa(p.width_mbs, p.height_mbs, n_pics,
                  np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.uint8),
                  np.zeros(n, np.int8), np.zeros(n, np.uint8),
                  np.zeros((n, 16), np.uint8), np.zeros((n, 384), np.int16),
                  cb_qp_offset=p.cb_qp_offset,
                  cr_qp_offset=p.cr_qp_offset if p.profile_idc >= 100 else p.cb_qp_offset)
        out.mb_kind = soa.mb_kind.ctypes.data
        out.i16_mode = soa.i16_mode.ctypes.data
        out.chroma_mode = soa.chroma_mode.ctypes.data
        out.cbp = soa.cbp.ctypes.data
        out.luma_modes = soa.luma_modes.ctypes.data
        out.qp_y = soa.qp_y.ctypes.data
        out.coeff = soa.coeff.ctypes.data
    if not lib.mvs_generate(C.byref(p), C.byref(out)):
        raise RuntimeError("mvs_generate failed")
    if soa is not None:
        soa.lists4x4 = np.ctypeslib.as_array(out.lists4x4).copy().reshape(6, 16)
        soa.lists8x8 = np.ctypeslib.as_array(out.lists8x8).copy().reshape(2, 64)
    data = bytes(stream[:out.stream_len]) if want_stream else None
    return data, soa
