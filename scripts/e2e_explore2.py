"""Dev script: end-to-end (pinned packed SoA -> GPU -> pinned RGB) throughput: blocking call vs two submissions in flight,
chunk sizes, full size and 1/4-size thumbnails.   python scripts/e2e_explore2.py [E]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from minivideo_b200 import api, synth

E = int(sys.argv[1]) if len(sys.argv) > 1 else 384
G = 16
_, soa = synth.generate(G, "1080p", want_stream=False, seed=0xC0FFEE + 2)
W, H = soa.width, soa.height
ctx = api.Context(0, soa.width_mbs, soa.height_mbs, 1000)
ctx.set_sps_from(soa)
packed = api.Packed(soa, n_pics=E, pinned=True)
for scale in (1, 4):
    px = (W // scale) * (H // scale) * 3
    outs = [api.PinnedArray((E, px), np.uint8), api.PinnedArray((E, px), np.uint8)]
    for chunk in (0, 24, 48, 96, 128, 192):
        ctx.set_pipeline(chunk)
        ctx.decode_host_packed(packed, None, outs[0].array, scale)
        K = 4
        t = time.perf_counter()
        for _ in range(K):
            ctx.decode_host_packed(packed, None, outs[0].array, scale)
        blocking = E * K / (time.perf_counter() - t)
        t = time.perf_counter()
        tk = api.submit_packed(ctx, packed, None, outs[0].array, scale)
        for i in range(1, K):
            t2 = api.submit_packed(ctx, packed, None, outs[i & 1].array, scale)
            ctx.wait(tk)
            tk = t2
        ctx.wait(tk)
        pipelined = E * K / (time.perf_counter() - t)
        print(f"scale {scale} E={E} chunk={chunk:3d}: blocking {blocking:8.0f}  two in flight {pipelined:8.0f} pictures/s"
              f"  (D2H {pipelined * px / 1e9:5.1f} GB/s, H2D {pipelined * packed.nbytes / E / 1e9:5.1f} GB/s)", flush=True)
    del outs
ctx.close()
