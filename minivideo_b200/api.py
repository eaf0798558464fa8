"""ctypes mirror of include/mvgpu.h (libmvgpu.so).

This module only binds the C ABI; it never computes pixels itself.  If the CUDA
library is missing, or there is no GPU, construction fails loudly -- there is
no CPU fallback on the product path.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

MVG_SUCCESS, MVG_FAILURE, MVG_UNSUPPORTED = 1, 0, -1
LIB_PATH = Path(__file__).resolve().parent / "libmvgpu.so"

# every symbol include/mvgpu.h declares
EXPORTS = (
    "mvg_create", "mvg_destroy", "mvg_last_error", "mvg_set_sps", "mvg_build_level_scale",
    "mvg_upload", "mvg_clone_slot", "mvg_run", "mvg_run_rgb", "mvg_run_thumbs", "mvg_set_pipeline_mode", "mvg_sync", "mvg_get_timing", "mvg_mark", "mvg_mark_elapsed",
    "mvg_download_yuv420", "mvg_download_rgb", "mvg_download_residual", "mvg_decode_host",
    "mvg_pack_batch", "mvg_decode_host_packed", "mvg_submit", "mvg_submit_packed", "mvg_wait", "mvg_poll",
    "mvg_device_count", "mvg_set_pipeline", "mvg_host_alloc", "mvg_host_free", "mvg_width", "mvg_height", "mvg_max_pics", "mvg_sm_count",
)


class MvgError(RuntimeError):
    pass


class PackedBatch(C.Structure):
    """mvg_packed_batch of include/mvgpu.h."""
    _fields_ = [("n_pics", C.c_int32), ("mb_kind", C.c_void_p), ("i16_mode", C.c_void_p),
                ("chroma_mode", C.c_void_p), ("qp_y", C.c_void_p), ("luma_modes", C.c_void_p),
                ("nz_blocks", C.c_void_p), ("word_off", C.c_void_p), ("pic_off", C.c_void_p), ("words", C.c_void_p)]


WORDS_PER_MB = 408      # MVG_PACKED_WORDS_PER_MB


class Batch(C.Structure):
    _fields_ = [("n_pics", C.c_int32), ("mb_kind", C.c_void_p), ("i16_mode", C.c_void_p),
                ("chroma_mode", C.c_void_p), ("qp_y", C.c_void_p), ("cbp", C.c_void_p),
                ("luma_modes", C.c_void_p), ("coeff", C.c_void_p)]


class Timing(C.Structure):
    _fields_ = [("k1_dequant_idct_ms", C.c_float), ("k2_wavefront_ms", C.c_float),
                ("k3_rgb_ms", C.c_float), ("total_ms", C.c_float), ("launches", C.c_int32), ("fused_ms", C.c_float)]


PIPELINE_FUSED, PIPELINE_SPLIT = 0, 1


_LIB = None


def load_library() -> C.CDLL:
    """Load libmvgpu.so (built in-tree by minivideo_b200.build).  Raises if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not LIB_PATH.exists():
        raise MvgError(f"{LIB_PATH} is missing: run `python -m minivideo_b200.build` "
                       "(the CUDA extension is required; there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    vp, i32 = C.c_void_p, C.c_int
    lib.mvg_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32]
    lib.mvg_destroy.argtypes = [vp]
    lib.mvg_last_error.argtypes = [vp]
    lib.mvg_last_error.restype = C.c_char_p
    lib.mvg_set_sps.argtypes = [vp, i32, i32, vp, vp, i32, i32]
    lib.mvg_build_level_scale.argtypes = [vp, vp, vp, vp]
    lib.mvg_upload.argtypes = [vp, C.POINTER(Batch), i32]
    lib.mvg_clone_slot.argtypes = [vp, i32, i32]
    lib.mvg_run.argtypes = [vp, i32, i32, i32]
    lib.mvg_run_rgb.argtypes = [vp, i32, i32]
    lib.mvg_run_thumbs.argtypes = [vp, i32, i32, i32]
    lib.mvg_set_pipeline_mode.argtypes = [vp, i32]
    lib.mvg_sync.argtypes = [vp]
    lib.mvg_get_timing.argtypes = [vp, C.POINTER(Timing)]
    lib.mvg_mark.argtypes = [vp, i32]
    lib.mvg_mark_elapsed.argtypes = [vp, C.POINTER(C.c_float)]
    lib.mvg_download_yuv420.argtypes = [vp, i32, vp, vp, vp]
    lib.mvg_download_rgb.argtypes = [vp, i32, vp]
    lib.mvg_download_residual.argtypes = [vp, i32, vp]
    lib.mvg_decode_host.argtypes = [vp, C.POINTER(Batch), vp, vp, i32]
    lib.mvg_pack_batch.argtypes = [vp, i32, i32, vp, vp, vp, vp, C.c_size_t, i32]
    lib.mvg_decode_host_packed.argtypes = [vp, C.POINTER(PackedBatch), vp, vp, i32]
    lib.mvg_submit.argtypes = [vp, C.POINTER(Batch), vp, vp, i32, C.POINTER(C.c_int32)]
    lib.mvg_submit_packed.argtypes = [vp, C.POINTER(PackedBatch), vp, vp, i32, C.POINTER(C.c_int32)]
    lib.mvg_wait.argtypes = [vp, i32]
    lib.mvg_poll.argtypes = [vp, i32, C.POINTER(C.c_int)]
    lib.mvg_set_pipeline.argtypes = [vp, i32]
    lib.mvg_host_alloc.argtypes = [C.c_size_t]
    lib.mvg_host_alloc.restype = vp
    lib.mvg_host_free.argtypes = [vp]
    for name in ("mvg_width", "mvg_height", "mvg_max_pics", "mvg_sm_count"):
        getattr(lib, name).argtypes = [vp]
    _LIB = lib
    return lib


def build_level_scale(lists4x4: np.ndarray | None, list8x8: np.ndarray | None):
    """(ls4[3,6,16], ls8[6,64]) int32 from zig-zag scaling lists; None = flat 16."""
    lib = load_library()
    ls4 = np.zeros((3, 6, 16), np.int32)
    ls8 = np.zeros((6, 64), np.int32)
    l4 = np.ascontiguousarray(lists4x4[:3], np.uint8) if lists4x4 is not None else None
    l8 = np.ascontiguousarray(list8x8, np.uint8) if list8x8 is not None else None
    rc = lib.mvg_build_level_scale(l4.ctypes.data if l4 is not None else None,
                                   l8.ctypes.data if l8 is not None else None,
                                   ls4.ctypes.data, ls8.ctypes.data)
    if rc != MVG_SUCCESS:
        raise MvgError("mvg_build_level_scale failed")
    return ls4, ls8


class PinnedArray:
    """A numpy view over cudaHostAlloc'ed memory (freed with the object)."""

    def __init__(self, shape, dtype):
        lib = load_library()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = lib.mvg_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise MvgError(f"mvg_host_alloc({self.nbytes}) failed")
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                load_library().mvg_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def _batch_of(soa, keep: list) -> Batch:
    def arr(a, dt):
        a = np.ascontiguousarray(a, dt)
        keep.append(a)
        return a.ctypes.data
    b = Batch()
    b.n_pics = soa.n_pics
    b.mb_kind = arr(soa.mb_kind, np.uint8)
    b.i16_mode = arr(soa.i16_mode, np.uint8)
    b.chroma_mode = arr(soa.chroma_mode, np.uint8)
    b.qp_y = arr(soa.qp_y, np.int8)
    b.cbp = arr(soa.cbp, np.uint8)
    b.luma_modes = arr(soa.luma_modes, np.uint8)
    b.coeff = arr(soa.coeff, np.int16)
    return b


class Context:
    """One GPU reconstruction context (mvg_ctx).  Not thread-safe; one per GPU."""

    def __init__(self, device: int, max_w_mbs: int, max_h_mbs: int, max_pics: int):
        self.lib = load_library()
        self.handle = C.c_void_p()
        rc = self.lib.mvg_create(C.byref(self.handle), device, max_w_mbs, max_h_mbs, max_pics)
        if rc != MVG_SUCCESS:
            raise MvgError(self.lib.mvg_last_error(None).decode())
        self.max_pics = max_pics

    def close(self):
        if self.handle:
            self.lib.mvg_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != MVG_SUCCESS:
            raise MvgError(self.lib.mvg_last_error(self.handle).decode())

    # -- tables / geometry ---------------------------------------------------
    def set_sps(self, width_mbs, height_mbs, ls4, ls8, cb_qp_offset=0, cr_qp_offset=0):
        ls4 = np.ascontiguousarray(ls4, np.int32)
        ls8 = np.ascontiguousarray(ls8, np.int32)
        assert ls4.size == 288 and ls8.size == 384
        self._ck(self.lib.mvg_set_sps(self.handle, width_mbs, height_mbs, ls4.ctypes.data, ls8.ctypes.data,
                                      cb_qp_offset, cr_qp_offset))
        self.width_mbs, self.height_mbs = width_mbs, height_mbs

    def set_sps_from(self, soa):
        ls4, ls8 = build_level_scale(soa.lists4x4, soa.lists8x8[0])
        self.set_sps(soa.width_mbs, soa.height_mbs, ls4, ls8, soa.cb_qp_offset, soa.cr_qp_offset)

    @property
    def width(self):
        return self.lib.mvg_width(self.handle)

    @property
    def height(self):
        return self.lib.mvg_height(self.handle)

    @property
    def sm_count(self):
        return self.lib.mvg_sm_count(self.handle)

    # -- resident path ---------------------------------------------------------
    def upload(self, soa, first_slot=0):
        keep = []
        b = _batch_of(soa, keep)
        self._ck(self.lib.mvg_upload(self.handle, C.byref(b), first_slot))

    def clone_slot(self, src, dst):
        self._ck(self.lib.mvg_clone_slot(self.handle, src, dst))

    def run(self, first_slot, n_pics, rgb_scale=1):
        self._ck(self.lib.mvg_run(self.handle, first_slot, n_pics, rgb_scale))

    def run_rgb(self, first_slot, n_pics):
        """One fused kernel: levels -> full-size RGB24 (no tiles; download_yuv420 is not available afterwards)."""
        self._ck(self.lib.mvg_run_rgb(self.handle, first_slot, n_pics))

    def run_thumbs(self, first_slot, n_pics, rgb_scale):
        """One fused kernel for rgb_scale 2, 4, 8, 16: levels -> RGB24 at 1/rgb_scale size (no tiles)."""
        self._ck(self.lib.mvg_run_thumbs(self.handle, first_slot, n_pics, rgb_scale))

    def set_pipeline_mode(self, mode: int):
        self._ck(self.lib.mvg_set_pipeline_mode(self.handle, mode))

    def sync(self):
        self._ck(self.lib.mvg_sync(self.handle))

    def timing(self) -> Timing:
        t = Timing()
        self._ck(self.lib.mvg_get_timing(self.handle, C.byref(t)))
        return t

    def mark(self, which: int):
        self._ck(self.lib.mvg_mark(self.handle, which))

    def mark_elapsed_ms(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.mvg_mark_elapsed(self.handle, C.byref(ms)))
        return float(ms.value)

    def download_yuv420(self, slot) -> np.ndarray:
        w, h = self.width, self.height
        out = np.empty(w * h * 3 // 2, np.uint8)
        base = out.ctypes.data
        self._ck(self.lib.mvg_download_yuv420(self.handle, slot, base, base + w * h, base + w * h * 5 // 4))
        return out

    def download_rgb(self, slot, scale=1) -> np.ndarray:
        w, h = self.width // scale, self.height // scale
        out = np.empty((h, w, 3), np.uint8)
        self._ck(self.lib.mvg_download_rgb(self.handle, slot, out.ctypes.data))
        return out

    def download_residual(self, slot) -> np.ndarray:
        out = np.empty((self.width_mbs * self.height_mbs, 384), np.int16)
        self._ck(self.lib.mvg_download_residual(self.handle, slot, out.ctypes.data))
        return out

    def submit(self, soa, yuv_out, rgb_out, rgb_scale=1, keep: list | None = None) -> int:
        """mvg_submit(): returns the ticket.  `keep` receives the arrays that must stay alive until wait()."""
        keep = keep if keep is not None else []
        b = _batch_of(soa, keep)
        self._keep_alive = getattr(self, "_keep_alive", [])
        self._keep_alive.append(keep)
        t = C.c_int32(-1)
        self._ck(self.lib.mvg_submit(self.handle, C.byref(b), _ptr(yuv_out), _ptr(rgb_out), rgb_scale, C.byref(t)))
        return int(t.value)

    def wait(self, ticket: int):
        self._ck(self.lib.mvg_wait(self.handle, ticket))

    def poll(self, ticket: int) -> bool:
        done = C.c_int(0)
        self._ck(self.lib.mvg_poll(self.handle, ticket, C.byref(done)))
        return bool(done.value)

    def set_pipeline(self, chunk_pics: int):
        self._ck(self.lib.mvg_set_pipeline(self.handle, chunk_pics))

    # -- end-to-end path -------------------------------------------------------
    def decode_host(self, soa, yuv_out: np.ndarray | None, rgb_out: np.ndarray | None, rgb_scale=1,
                    batch: Batch | None = None):
        keep = []
        b = batch if batch is not None else _batch_of(soa, keep)
        self._ck(self.lib.mvg_decode_host(self.handle, C.byref(b),
                                          yuv_out.ctypes.data if yuv_out is not None else None,
                                          rgb_out.ctypes.data if rgb_out is not None else None, rgb_scale))

    def decode_host_packed(self, packed: "Packed", yuv_out: np.ndarray | None, rgb_out: np.ndarray | None, rgb_scale=1):
        self._ck(self.lib.mvg_decode_host_packed(self.handle, C.byref(packed.struct),
                                                 yuv_out.ctypes.data if yuv_out is not None else None,
                                                 rgb_out.ctypes.data if rgb_out is not None else None, rgb_scale))


def _ptr(a):
    return a.ctypes.data if a is not None else None


def submit_packed(ctx: "Context", packed: "Packed", yuv_out, rgb_out, rgb_scale=1) -> int:
    """mvg_submit_packed(): returns the ticket; ctx.wait(ticket) completes it."""
    t = C.c_int32(-1)
    ctx._ck(ctx.lib.mvg_submit_packed(ctx.handle, C.byref(packed.struct), _ptr(yuv_out), _ptr(rgb_out), rgb_scale, C.byref(t)))
    return int(t.value)


class Packed:
    """A batch in the packed transfer format (mvg_packed_batch), built from dense levels with mvg_pack_batch().
    With pinned=True every array lives in cudaHostAlloc'ed memory (full PCIe speed)."""

    def __init__(self, soa, n_pics=None, pinned=False, n_threads=0):
        lib = load_library()
        P = soa.n_pics if n_pics is None else n_pics
        N = soa.n_mbs
        reps = -(-P // soa.n_pics)
        self._keep = []

        def arr(shape, dt):
            if pinned:
                pa = PinnedArray(shape, dt)
                self._keep.append(pa)
                return pa.array
            return np.empty(shape, dt)

        def tiled(src):
            out = arr((P * N,) + src.shape[1:], src.dtype)
            out[...] = np.concatenate([src] * reps)[: P * N]
            return out
        self.n_pics, self.n_mbs = P, N
        self.mb_kind, self.i16_mode, self.chroma_mode = tiled(soa.mb_kind), tiled(soa.i16_mode), tiled(soa.chroma_mode)
        self.qp_y, self.luma_modes = tiled(soa.qp_y), tiled(soa.luma_modes)
        coeff = np.ascontiguousarray(np.concatenate([soa.coeff] * reps)[: P * N])
        self.nz_blocks, self.word_off = arr((P * N,), np.uint32), arr((P * N,), np.uint32)
        self.pic_off = arr((P + 1,), np.uint64)
        words = np.empty(P * N * WORDS_PER_MB, np.uint16)
        rc = lib.mvg_pack_batch(coeff.ctypes.data, P, N, self.nz_blocks.ctypes.data, self.word_off.ctypes.data,
                                self.pic_off.ctypes.data, words.ctypes.data, words.size, n_threads)
        if rc != 1:
            raise MvgError("mvg_pack_batch failed")
        n_words = int(self.pic_off[P])
        self.words = arr((max(n_words, 1),), np.uint16)
        self.words[:n_words] = words[:n_words]
        self.n_words = n_words
        self.struct = PackedBatch(P, *(a.ctypes.data for a in (self.mb_kind, self.i16_mode, self.chroma_mode, self.qp_y,
                                                               self.luma_modes, self.nz_blocks, self.word_off,
                                                               self.pic_off, self.words)))

    @property
    def nbytes(self) -> int:
        """Bytes that cross the bus for this batch."""
        return (self.mb_kind.nbytes + self.i16_mode.nbytes + self.chroma_mode.nbytes + self.qp_y.nbytes +
                self.luma_modes.nbytes + self.nz_blocks.nbytes + self.word_off.nbytes + 8 * self.n_pics + 2 * self.n_words)


def unpack_levels(pk: "Packed") -> np.ndarray:
    """Numpy restatement of the packed format (tests): packed -> dense levels [P*N, 384]."""
    out = np.zeros((pk.n_pics * pk.n_mbs, 384), np.int16)
    for p in range(pk.n_pics):
        base = int(pk.pic_off[p])
        for m in range(pk.n_mbs):
            mb = p * pk.n_mbs + m
            nzb = int(pk.nz_blocks[mb])
            w = base + int(pk.word_off[mb])
            blocks = [b for b in range(24) if (nzb >> b) & 1]
            lv = w + len(blocks)
            for i, b in enumerate(blocks):
                mask = int(pk.words[w + i])
                for k in range(16):
                    if (mask >> k) & 1:
                        out[mb, b * 16 + k] = pk.words[lv:lv + 1].view(np.int16)[0]
                        lv += 1
    return out


def reconstruct(soa, device=0, rgb_scale=1, want_residual=False, mode=PIPELINE_FUSED):
    """Convenience for tests: run a whole Soa through the resident path.
    Returns dict(yuv [P, 1.5WH], rgb [P, H/s, W/s, 3] | None, rgb_k3 (the same through the tiles + kernel 3),
    residual | None).  With the fused pipeline and rgb_scale 1, `rgb` comes from mvg_run_rgb() (one kernel, levels ->
    RGB24) and `rgb_k3` from mvg_run() (fused kernel -> tiles -> kernel 3); with rgb_scale 2, 4, 8, 16 `rgb` comes from
    mvg_run_thumbs() (one kernel, levels -> thumbnail); otherwise both are the same array."""
    ctx = Context(device, soa.width_mbs, soa.height_mbs, soa.n_pics)
    try:
        ctx.set_pipeline_mode(mode)
        ctx.set_sps_from(soa)
        ctx.upload(soa, 0)
        ctx.run(0, soa.n_pics, rgb_scale)
        ctx.sync()
        timing = ctx.timing()
        yuv = np.stack([ctx.download_yuv420(i) for i in range(soa.n_pics)])
        rgb = np.stack([ctx.download_rgb(i, rgb_scale) for i in range(soa.n_pics)]) if rgb_scale >= 1 else None
        res = np.concatenate([ctx.download_residual(i) for i in range(soa.n_pics)]) if want_residual else None
        rgb_k3 = rgb
        if rgb_scale == 1 and mode == PIPELINE_FUSED:
            ctx.run_rgb(0, soa.n_pics)
            ctx.sync()
            rgb = np.stack([ctx.download_rgb(i, 1) for i in range(soa.n_pics)])
        elif rgb_scale in (2, 4, 8, 16) and mode == PIPELINE_FUSED:
            ctx.run_thumbs(0, soa.n_pics, rgb_scale)        # one kernel: levels -> thumbnail
            ctx.sync()
            rgb = np.stack([ctx.download_rgb(i, rgb_scale) for i in range(soa.n_pics)])
        return dict(yuv=yuv, rgb=rgb, rgb_k3=rgb_k3, residual=res, timing=timing)
    finally:
        ctx.close()
