"""Dev script: first GPU parity check of the CUDA path against the CPU oracle."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from minivideo_b200 import api, synth
from oracle import cpu

def check(tag, n, scale=1, **kw):
    _, soa = synth.generate(n, want_stream=False, **kw)
    yuv_ref, res_ref = cpu.reconstruct(soa, want_residual=True)
    out = api.reconstruct(soa, rgb_scale=scale, want_residual=True)
    W, H = soa.width, soa.height
    rgb_ref = cpu.yuv_to_rgb(yuv_ref, W, H, scale)
    dres = int((out["residual"] != res_ref).sum())
    dy = out["yuv"] != yuv_ref
    drgb = int((out["rgb"] != rgb_ref).sum())
    t = out["timing"]
    print(f"{tag:14s} resid_mism={dres} yuv_mism={int(dy.sum())} rgb_mism={drgb} "
          f"k1={t.k1_dequant_idct_ms:.3f} k2={t.k2_wavefront_ms:.3f} k3={t.k3_rgb_ms:.3f} ms", flush=True)
    if dy.sum():
        p, off = np.argwhere(dy)[0]
        if off < W * H:
            print("   first luma mismatch pic", p, "x", off % W, "y", off // W, "mb", (off // W // 16) * soa.width_mbs + (off % W) // 16,
                  "kind", soa.mb_kind[p * soa.n_mbs + (off // W // 16) * soa.width_mbs + (off % W) // 16])
        else:
            print("   first chroma mismatch pic", p, "off", off - W * H)
    return dres == 0 and dy.sum() == 0 and drgb == 0

ok = True
for k in range(3):
    ok &= check(f"kind{k}", 2, width_mbs=6, height_mbs=5, profile_idc=100, transform8x8=1, scaling_lists=1, force_kind=k, seed=100 + k, qp_min=10, qp_max=45, init_qp=30)
for m in range(9):
    ok &= check(f"mode{m}", 1, width_mbs=6, height_mbs=6, profile_idc=100, transform8x8=1, force_mode=m, seed=200 + m)
ok &= check("1x1", 3, width_mbs=1, height_mbs=1, profile_idc=66, seed=9)
ok &= check("1xN", 2, width_mbs=1, height_mbs=9, profile_idc=100, transform8x8=1, seed=10)
ok &= check("Nx1", 2, width_mbs=13, height_mbs=1, profile_idc=100, transform8x8=1, seed=11)
ok &= check("cif", 4, config="cif")
ok &= check("cif_s4", 2, scale=4, config="cif")
ok &= check("720p", 3, config="720p")
ok &= check("1080p", 8, config="1080p")
ok &= check("1080p_s4", 2, scale=4, config="1080p")
print("ALL OK" if ok else "MISMATCHES")
sys.exit(0 if ok else 1)
