"""Development aid: time the fused kernel in its two output modes (and the split kernels) over F resident 1080p pictures.
    python tests/tools/kf_time.py [F] [reps] [quick]
Prints a SHA-1 of what the two modes produce for three slots (RGB24, planar YUV), so that builds with different
-D switches (scripts/try_kf.sh) can be compared with each other at a glance; `quick` skips the split pipeline."""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from minivideo_b200 import api, synth  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
quick = len(sys.argv) > 3
G = 16
_, soa = synth.generate(G, "1080p", want_stream=False, seed=0xC0FFEE + 2)
ctx = api.Context(0, soa.width_mbs, soa.height_mbs, F)
ctx.set_sps_from(soa)
ctx.upload(soa, 0)
for s in range(G, F):
    ctx.clone_slot(s % G, s)
ctx.sync()


def timed(fn, name, digest):
    for _ in range(3):
        fn()
    ctx.sync()
    ts = []
    for _ in range(reps):
        fn()
        t = ctx.timing()
        ts.append((t.fused_ms, t.k1_dequant_idct_ms, t.k2_wavefront_ms, t.k3_rgb_ms, t.total_ms))
    m = np.mean(ts, axis=0)
    h = hashlib.sha1()
    for slot in (0, G + 5, F - 1):
        h.update(digest(slot).tobytes())
    print(f"{name:28s} fused {m[0]:7.3f}  k1 {m[1]:6.3f}  k2 {m[2]:6.3f}  k3 {m[3]:6.3f}  total {m[4]:7.3f} ms  -> {F / m[4]:7.1f} k pictures/s  sha1 {h.hexdigest()[:12]}", flush=True)


timed(lambda: ctx.run_rgb(0, F), "fused -> RGB24", lambda s: ctx.download_rgb(s, 1))
timed(lambda: ctx.run(0, F, 0), "fused -> tiles", lambda s: ctx.download_yuv420(s))
timed(lambda: ctx.run_thumbs(0, F, 4), "fused -> RGB24 at 1/4", lambda s: ctx.download_rgb(s, 4))
if not quick:
    timed(lambda: ctx.run(0, F, 1), "fused -> tiles -> k3", lambda s: ctx.download_rgb(s, 1))
    timed(lambda: ctx.run(0, F, 4), "fused -> tiles -> k3 at 1/4", lambda s: ctx.download_rgb(s, 4))
    ctx.set_pipeline_mode(api.PIPELINE_SPLIT)
    timed(lambda: ctx.run(0, F, 1), "split k1 k2 k3", lambda s: ctx.download_rgb(s, 1))
ctx.close()
