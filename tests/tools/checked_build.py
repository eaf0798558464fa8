"""Memory-safety evidence without compute-sanitizer (closed on this pool, see profiles/r02_sanitizer_closed.txt).

    python tests/tools/checked_build.py            (on a GPU box; rebuilds libmvgpu.so with -DMVG_CHECKED, restores it afterwards)

What runs: the small parity cases that exercise every edge of the kernels (one macroblock, one row, one column, two
columns, CIF, a High-profile picture with 8x8 blocks and scaling lists, a 1080p picture) plus batches of random bytes in
every SoA field, through the resident entry points, both pipelines, and the end-to-end entry points, with
  * MVG_DEBUG_GUARD=1: every device allocation between two 64 KB guard bands (checked after every case: a write outside a
    buffer) and poisoned instead of zeroed (every case must still be bit-exact: nothing reads what it did not write);
    the whole sequence runs twice with the cases in a different order, so that what a buffer holds from the previous
    case differs;
  * -DMVG_CHECKED kernels: every computed shared-memory address of the prediction stage and every list index of the
    transform stage is compared with the bounds of the warp's record, records carry canary words between their members.
Prints one line per case and a summary; exit status 0 only if every counter is zero and every case is bit-exact."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
os.environ["MVG_DEBUG_GUARD"] = "1"
os.environ["MVG_EXTRA_DEFINES"] = (os.environ.get("MVG_EXTRA_DEFINES", "") + " MVG_CHECKED").strip()
from minivideo_b200 import build  # noqa: E402

build.build_gpu(force=True)
from minivideo_b200 import api, synth  # noqa: E402
from oracle import cpu  # noqa: E402

CASES = [
    ("one macroblock", 3, dict(width_mbs=1, height_mbs=1, profile_idc=100, transform8x8=1, seed=66)),
    ("one row", 2, dict(width_mbs=33, height_mbs=1, profile_idc=100, transform8x8=1, seed=67)),
    ("one column", 2, dict(width_mbs=1, height_mbs=33, profile_idc=100, transform8x8=1, seed=68)),
    ("two columns", 2, dict(width_mbs=2, height_mbs=17, profile_idc=100, transform8x8=1, seed=69)),
    ("three columns", 2, dict(width_mbs=3, height_mbs=5, profile_idc=100, transform8x8=1, scaling_lists=1, seed=70)),
    ("CIF baseline", 2, dict(config="cif")),
    ("High 10x6, lists, offsets", 2, dict(width_mbs=10, height_mbs=6, profile_idc=100, transform8x8=1, scaling_lists=1, cb_qp_offset=2, cr_qp_offset=-3, seed=77)),
    ("hostile levels", 1, dict(width_mbs=8, height_mbs=8, profile_idc=100, transform8x8=1, scaling_lists=1, seed=65, level_scale_x10=30000, max_level=8191, qp_min=40, qp_max=51, init_qp=45, luma_cbp_percent=100)),
    ("1080p", 2, dict(config="1080p", seed=62)),
]
lib = api.load_library()
lib.mvg_debug_check.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
lib.mvg_debug_check.restype = C.c_longlong
lib.mvg_debug_check_kernels.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
failures = 0


def audit(ctx, what):
    global failures
    rep = C.create_string_buffer(1024)
    bad = lib.mvg_debug_check(ctx.handle, rep, 1024)
    cnt = (C.c_ulonglong * 8)()
    rc = lib.mvg_debug_check_kernels(ctx.handle, cnt)
    k = list(cnt)[:4]
    ok = bad == 0 and rc == 1 and not any(k)
    failures += not ok
    print(f"  {what:44s} guard bytes damaged {bad} {rep.value.decode()}  kernel checks [address, list, canary, table] = {k}  {'ok' if ok else 'FAIL'}", flush=True)


def run_case(name, n, kw):
    global failures
    _, soa = synth.generate(n, want_stream=False, **kw)
    want = cpu.reconstruct(soa)[0]
    want_rgb = cpu.yuv_to_rgb(want, soa.width, soa.height, 1)
    print(f"{name}: {n} pictures {soa.width}x{soa.height}")
    for mode, label in ((api.PIPELINE_FUSED, "fused"), (api.PIPELINE_SPLIT, "split")):
        ctx = api.Context(0, soa.width_mbs, soa.height_mbs, max(3, n))
        ctx.set_pipeline_mode(mode)
        ctx.set_sps_from(soa)
        ctx.upload(soa, 0)
        ctx.run(0, n, 1); ctx.sync()
        exact = all(np.array_equal(ctx.download_yuv420(i), want[i]) and np.array_equal(ctx.download_rgb(i), want_rgb[i]) for i in range(n))
        if mode == api.PIPELINE_FUSED:
            ctx.run_rgb(0, n); ctx.sync()
            exact &= all(np.array_equal(ctx.download_rgb(i), want_rgb[i]) for i in range(n))
            if soa.width % 4 == 0 and soa.height % 4 == 0:
                ctx.run(0, n, 4); ctx.sync()
                exact &= np.array_equal(ctx.download_rgb(0, 4), cpu.yuv_to_rgb(want[:1], soa.width, soa.height, 4)[0])
        yuv = np.zeros_like(want); rgb = np.zeros((n, soa.width * soa.height * 3), np.uint8)
        ctx.decode_host_packed(api.Packed(soa), yuv, rgb, 1)
        exact &= np.array_equal(yuv, want) and np.array_equal(rgb.reshape(want_rgb.shape), want_rgb)
        rgb[...] = 0
        ctx.decode_host(soa, None, rgb, 1)
        exact &= np.array_equal(rgb.reshape(want_rgb.shape), want_rgb)
        failures += not exact
        audit(ctx, f"{label}: resident, RGB-only, 1/4 size, end to end{'' if exact else '  NOT BIT-EXACT'}")
        ctx.close()


def run_garbage(seed):
    """random bytes in every field: the kernels must finish inside their buffers whatever the side information says"""
    rng = np.random.default_rng(seed)
    _, soa = synth.generate(2, want_stream=False, width_mbs=7, height_mbs=5, profile_idc=100, transform8x8=1, seed=5)
    for f in ("mb_kind", "i16_mode", "chroma_mode", "qp_y", "cbp", "luma_modes", "coeff"):
        a = getattr(soa, f)
        setattr(soa, f, rng.integers(np.iinfo(a.dtype).min, np.iinfo(a.dtype).max, a.shape, dtype=a.dtype, endpoint=True))
    for mode, label in ((api.PIPELINE_FUSED, "fused"), (api.PIPELINE_SPLIT, "split")):
        ctx = api.Context(0, soa.width_mbs, soa.height_mbs, 3)
        ctx.set_pipeline_mode(mode)
        ctx.set_sps_from(soa)
        ctx.upload(soa, 0)
        ctx.run(0, 2, 1); ctx.sync()
        if mode == api.PIPELINE_FUSED:
            ctx.run_rgb(0, 2); ctx.sync()
        audit(ctx, f"{label}: random bytes in every SoA field (seed {seed})")
        ctx.close()


for order in (CASES, CASES[::-1]):
    for name, n, kw in order:
        run_case(name, n, kw)
    for seed in (1, 2, 3):
        run_garbage(seed)
print("SUMMARY:", "all cases bit-exact, no guard byte damaged, no kernel check failed" if failures == 0 else f"{failures} FAILURES")
del os.environ["MVG_EXTRA_DEFINES"]
build.build_gpu(force=True)         # back to the product build
sys.exit(1 if failures else 0)
