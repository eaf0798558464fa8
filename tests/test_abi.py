"""C-ABI checks that need no GPU: the library loads, exports every symbol include/mvgpu.h
declares, refuses to run without a device, and its host-side table builders are right."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols(header: Path):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(mv[gsf]_[a-z0-9_]+)\s*\(", text)))


def test_libmvgpu_exports_every_declared_symbol():
    from minivideo_b200 import api
    lib = api.load_library()
    syms = declared_symbols(ROOT / "include" / "mvgpu.h")
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in mvgpu.h but not exported"
    assert set(syms) == set(api.EXPORTS)


def test_libmvsynth_exports_every_declared_symbol():
    from minivideo_b200 import synth
    lib = synth._lib()
    for s in declared_symbols(ROOT / "include" / "mvsynth.h"):
        assert hasattr(lib, s)


def test_no_device_means_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from minivideo_b200 import api
    with pytest.raises(api.MvgError, match="no CPU fallback"):
        api.Context(0, 4, 4, 1)


def test_null_context_calls_fail_cleanly():
    from minivideo_b200 import api
    lib = api.load_library()
    assert lib.mvg_run(None, 0, 1, 1) == api.MVG_FAILURE
    assert lib.mvg_sync(None) == api.MVG_FAILURE
    assert lib.mvg_destroy(None) == api.MVG_FAILURE
    assert lib.mvg_last_error(None) is not None


def test_product_sources_never_touch_the_oracle():
    """The product path (package + C sources) must not import, link or call anything under oracle/."""
    for p in list((ROOT / "minivideo_b200").rglob("*.py")) + list((ROOT / "minivideo_b200" / "csrc").glob("*")):
        if p.suffix in (".py", ".c", ".cu", ".cuh", ".h") and p.name != "build.py":
            text = p.read_text()
            assert "recon_oracle" not in text and "from oracle" not in text and "import oracle" not in text, p


def test_pred_tap_tables_reproduce_the_oracle_predictors():
    """mvg_build_luts(): evaluate the 4-tap tables on random neighbours in numpy and compare with the
    oracle's spec-shaped predictors (through one-block pictures)."""
    from minivideo_b200 import api
    from minivideo_b200.synth import Soa
    from oracle import cpu

    class Luts(C.Structure):
        _fields_ = [("lut4", C.c_uint32 * (16 * 32)), ("lut8", C.c_uint32 * (16 * 32))]
    lib = api.load_library()
    luts = Luts()
    lib.mvg_build_luts(C.byref(luts))
    lut4 = np.frombuffer(luts.lut4, np.uint32).reshape(16, 32)
    lut8 = np.frombuffer(luts.lut8, np.uint32).reshape(16, 32)
    assert not lut8[9:].any()                               # rows of mode values no conforming stream has
    import re
    hdr = (ROOT / "minivideo_b200" / "csrc" / "mvg_internal.h").read_text()
    STRIDE = int(re.search(r"#define MVG_LT_STRIDE\s+(\d+)", hdr).group(1))
    BIAS = STRIDE + 1                                       # MVG_LUT4_BIAS
    taps4 = lambda row: np.stack([(lut4[row, :16] >> (8 * k)) & 255 for k in range(4)], -1).astype(np.int64) - BIAS
    assert np.array_equal(lut4[:, :16], lut4[:, 16:])       # both lane halves read the same taps
    # rows 11 / 15: modes 3 / 7 with the taps on p[4..7,-1] moved to p[3,-1]
    for row, mode in ((11, 3), (15, 7)):
        off = taps4(mode)
        top = (off + 1) // STRIDE == -1
        assert np.array_equal(taps4(row), np.where(top & (off + STRIDE > 3), -STRIDE + 3, off))
    rng = np.random.default_rng(1)

    # 2x2-MB picture: MBs 0,1,2 are I16x16 with random DC so that MB 3 sees random neighbours;
    # MB 3 is I4x4 (or I8x8) with zero residual => its samples ARE the prediction.
    for kind, n, lut in ((0, 4, None), (1, 8, None)):
        for mode in range(9):
            if mode == 2:
                continue
            soa = Soa(3, 2, 1, np.full(6, 2, np.uint8), np.zeros(6, np.uint8), np.zeros(6, np.uint8),
                      np.full(6, 30, np.int8), np.zeros(6, np.uint8), np.full((6, 16), 2, np.uint8),
                      np.zeros((6, 384), np.int16))
            soa.i16_mode[:] = 2
            soa.coeff[:, :256] = rng.integers(-6, 7, (6, 256))
            soa.coeff[:, ::16] = rng.integers(-40, 40, (6, 24))
            soa.mb_kind[4] = kind
            soa.coeff[4] = 0
            soa.luma_modes[4] = mode
            yuv, _ = cpu.reconstruct(soa)
            Y = yuv[0, :48 * 32].reshape(32, 48).astype(np.int32)
            # block 0 of MB 4 sits at (16,16); evaluate the tap table for it
            x0, y0 = 16, 16
            if kind == 0:
                pred = np.zeros((4, 4), np.int32)
                for y in range(4):
                    for x in range(4):
                        s = 0
                        for k in range(4):
                            off = int(taps4(mode)[y * 4 + x, k])            # relative to the block origin
                            dy = (off + 1) // STRIDE
                            dx = off - dy * STRIDE
                            s += Y[y0 + dy, x0 + dx]
                        pred[y, x] = (s + 2) >> 2
                assert np.array_equal(pred, Y[y0:y0 + 4, x0:x0 + 4]), (kind, mode)
            else:
                top = Y[y0 - 1, x0:x0 + 16]; left = Y[y0:y0 + 8, x0 - 1]; tl = Y[y0 - 1, x0 - 1]
                ft = np.zeros(16, np.int32); fl = np.zeros(8, np.int32)
                ft[0] = (tl + 2 * top[0] + top[1] + 2) >> 2
                for i in range(1, 15):
                    ft[i] = (top[i - 1] + 2 * top[i] + top[i + 1] + 2) >> 2
                ft[15] = (top[14] + 3 * top[15] + 2) >> 2
                fl[0] = (tl + 2 * left[0] + left[1] + 2) >> 2
                for i in range(1, 7):
                    fl[i] = (left[i - 1] + 2 * left[i] + left[i + 1] + 2) >> 2
                fl[7] = (left[6] + 3 * left[7] + 2) >> 2
                ftl = (top[0] + 2 * tl + left[0] + 2) >> 2
                line = np.concatenate([fl[::-1], [ftl], ft])
                nxt = np.concatenate([line[1:], line[-1:]]); prv = np.concatenate([line[:1], line[:-1]])
                variants = {0: line, 8: (line + nxt + 1) >> 1, 16: (prv + 2 * line + nxt + 2) >> 2}
                pred = np.zeros((8, 8), np.int32)
                for y in range(8):
                    for x in range(8):
                        lane = (y // 4) * 16 + (x // 4) * 8 + (y % 4) * 2 + (x % 4) // 2
                        e = (int(lut8[mode, lane]) >> (16 * (x % 2))) & 0xffff
                        idx, var = e // 4, e % 4                            # line entry idx, byte var of its word
                        pred[y, x] = variants[8 * var][idx]
                assert np.array_equal(pred, Y[y0:y0 + 8, x0:x0 + 8]), (kind, mode)
