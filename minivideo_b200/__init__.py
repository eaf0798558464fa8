"""B200-native H.264 intra-picture reconstruction (the MiniVideo thumbnailer hot path).

Only the path named in BASELINE.json lives here: `csrc/` holds the CUDA kernels,
the C ABI (include/mvgpu.h) and the host front end; the Python modules are thin
ctypes mirrors of those C interfaces used by tests and bench.py.
"""
__all__ = ["build", "synth"]
