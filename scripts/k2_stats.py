"""Dev script: in-kernel cycle accounting of kernel 2 (build with MVG_K2_PROFILE=1)."""
import ctypes as C, sys
sys.path.insert(0, ".")
from minivideo_b200 import api, synth
F, G = int(sys.argv[1]) if len(sys.argv) > 1 else 384, 8
_, soa = synth.generate(G, "1080p", want_stream=False)
ctx = api.Context(0, soa.width_mbs, soa.height_mbs, F)
ctx.set_sps_from(soa); ctx.upload(soa, 0)
for s in range(G, F): ctx.clone_slot(s % G, s)
lib = api.load_library()
lib.mvg_dev_k2_stats.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
out = (C.c_ulonglong * 16)()
for it in range(2):
    ctx.run(0, F, 1); ctx.sync()
    t = ctx.timing()
    lib.mvg_dev_k2_stats(ctx.handle, out)
    v = list(out)
    nmb = F * soa.n_mbs
    tot = v[8]
    names = ["row prologue", "halo+issue", "resid wait", "predict", "tail(write-out, hand-over)", "slow entries", "poll iters", "halo section"]
    print(f"k2 {t.k2_wavefront_ms:.3f} ms; warp-cycles total {tot/1e9:.2f}G  per MB {tot/nmb:.0f}")
    for n, x in zip(names, v[:8]):
        print(f"  {n:28s} {x/nmb:10.2f} per MB   {100*x/tot if tot else 0:5.1f}% of warp time")
