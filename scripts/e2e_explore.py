"""Dev script: end-to-end (pinned host -> GPU -> pinned host) throughput vs batch and chunk size."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from minivideo_b200 import api, synth

G = 8
_, soa = synth.generate(G, "1080p", want_stream=False)
N, W, H = soa.n_mbs, soa.width, soa.height
# raw PCIe
a = torch.empty(1 << 30, dtype=torch.uint8).pin_memory(); b = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
da = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); db = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for both in (False, True):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(4):
        with torch.cuda.stream(s1): da.copy_(a, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): b.copy_(db, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"PCIe H2D{'+D2H' if both else ''}: {4 * (1 << 30) / dt / 1e9:.1f} GB/s per direction", flush=True)
del a, b, da, db

for scale in (1, 4):
    for E in (96, 384):
        ctx = api.Context(0, soa.width_mbs, soa.height_mbs, 384)
        ctx.set_sps_from(soa)
        reps = -(-E // G)
        pin = {k: api.PinnedArray((E * N,) + getattr(soa, k).shape[1:], getattr(soa, k).dtype) for k in
               ("mb_kind", "i16_mode", "chroma_mode", "qp_y", "cbp", "luma_modes", "coeff")}
        for k, pa in pin.items():
            pa.array[...] = np.concatenate([getattr(soa, k)] * reps)[: E * N]
        out = api.PinnedArray((E, (W // scale) * (H // scale) * 3), np.uint8)
        batch = api.Batch(); batch.n_pics = E
        for k, pa in pin.items(): setattr(batch, k, pa.ptr)
        for chunk in (0, 4, 8, 16, 32, 64):
            ctx.set_pipeline(chunk)
            ctx.decode_host(None, None, out.array, scale, batch=batch)
            t = time.perf_counter()
            for _ in range(3): ctx.decode_host(None, None, out.array, scale, batch=batch)
            dt = (time.perf_counter() - t) / 3
            print(f"scale {scale} E={E:4d} chunk={chunk:3d}: {E / dt:8.0f} pictures/s  ({E * N * 789 / dt / 1e9:.1f} GB/s H2D)", flush=True)
        ctx.close(); del pin, out
