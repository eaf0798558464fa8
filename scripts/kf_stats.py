"""Dev script: wait accounting of the fused kernel (build with MVG_EXTRA_DEFINES=KF_STATS)."""
import ctypes as C, sys
sys.path.insert(0, ".")
from minivideo_b200 import api, synth
F, G = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 16
_, soa = synth.generate(G, "1080p", want_stream=False, seed=0xC0FFEE + 2)
ctx = api.Context(0, soa.width_mbs, soa.height_mbs, F)
ctx.set_sps_from(soa); ctx.upload(soa, 0)
for s in range(G, F): ctx.clone_slot(s % G, s)
lib = api.load_library()
lib.mvg_dev_k2_stats.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
out = (C.c_ulonglong * 16)()
for it in range(3):
    lib.mvg_dev_k2_stats(ctx.handle, out)       # clear
    ctx.run_rgb(0, F); ctx.sync()
    t = ctx.timing()
    lib.mvg_dev_k2_stats(ctx.handle, out)
    v = list(out)
    rows, nmb = v[5], F * soa.n_mbs
    print(f"kf {t.fused_ms:.3f} ms, rows {rows}: row time {v[6]/max(rows,1):.0f} cycles ({v[6]/max(rows,1)/soa.width_mbs:.0f} per MB); "
          f"row-start wait {v[1]/max(rows,1):.0f} cycles, {v[0]/max(rows,1):.1f} polls per row; "
          f"in-row: {v[2]/max(rows,1):.2f} catch-ups per row, {v[3]/max(rows,1):.1f} polls per row, {v[4]/max(rows,1):.0f} cycles per row", flush=True)
