#!/usr/bin/env python
"""Benchmark of the H.264 intra reconstruction hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--frames F] [--distinct G] [--rgb-scale S] [--e2e-frames E]

One "step" = one pass of the hot path (kernel 1 dequant/IDCT, kernel 2 wavefront
prediction, kernel 3 RGB) over one batch of F synthetic 1920x1088 High-profile
IDR pictures (BASELINE.json configs[2]) that are already resident in HBM.
Rank 0 prints ONE JSON line.  Multi-GPU: one process per GPU (torchrun), disjoint
pictures per GPU, no collective on the data path ("scaling": "weak").

--impl reference times the reference's own CPU implementation (oracle/_ref, the
unmodified reference compiled by oracle/Makefile) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "1080p_idr_frames_per_s_recon_rgb"
UNIT = "frames/s"
WORKLOAD = "configs[2]: 1920x1080 High-profile CAVLC, 8x8 transform + custom scaling lists, 1000 IDR frames on 1 B200"


# ----------------------------------------------------------------------------
# helpers

def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU; result(t0, t1) keeps the samples taken
    inside the timed region [t0, t1] (time.perf_counter())."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples = index, threading.Event(), []
        self.max_mhz, self.err = None, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons",
                                  getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None))
            while not self.stop_flag.is_set():
                mask = get_reasons(h) if get_reasons else 0
                self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mask))
                time.sleep(0.004)
        except Exception as e:  # pragma: no cover - depends on the box
            self.err = repr(e)

    def result(self, t0: float, t1: float):
        self.stop_flag.set()
        self.join(timeout=2)
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "sw_power_cap": 0x4, "hw_power_brake_slowdown": 0x80}
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "error": self.err}
        mask = 0
        for s in inside:
            mask |= s[2]
        return {"sm_mhz": float(np.median([s[1] for s in inside])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(k for k, bit in names.items() if mask & bit), "samples": len(inside)}


def algorithmic_bytes(n_mb: int, width: int, height: int, scale: int):
    """SURVEY.md section 8(d), per picture; kernel 2 reads the 16-byte control record
    kernel 1 writes instead of the 32 B/MB the survey budgeted (stated in DESIGN.md)."""
    meta_in = 21                      # mb_kind, i16_mode, chroma_mode, qp_y, cbp, 16 luma modes
    k1 = n_mb * (768 + meta_in) + n_mb * (768 + 16)
    k2 = n_mb * (768 + 16) + width * height * 3 // 2
    k3 = width * height * 3 // 2 + 3 * (width // scale) * (height // scale) if scale >= 1 else 0
    kf = n_mb * (768 + meta_in) + 3 * width * height      # fused: SoA in, RGB24 out
    return {"k1": k1, "k2": k2, "k3": k3, "kf": kf}


# ----------------------------------------------------------------------------
# reference arm: the unmodified reference decoder on the host cores

def reference_sample(n_procs: int, pics_per_proc: int, rgb: bool = True):
    """Every core decodes the same `pics_per_proc` 1080p pictures with the reference
    (CAVLC parse + reconstruction + mb_to_rgb, no file output).  Returns
    (frames_per_s, wall_s)."""
    from minivideo_b200 import synth
    from oracle import ref
    if not ref.available():
        raise RuntimeError("oracle/_ref/ref_decode is missing (run `make -C oracle ref` where the reference is mounted)")
    stream, _ = synth.generate(pics_per_proc, "1080p", want_soa=False)
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    with tempfile.TemporaryDirectory(dir=base, prefix="mvbench_") as d:
        path = os.path.join(d, "sample.264")
        Path(path).write_bytes(stream)
        ref.time_decode(path, 1, rgb=rgb, cwd=d)            # warm the page cache / binary
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=n_procs) as ex:
            secs = list(ex.map(lambda _: ref.time_decode(path, pics_per_proc, rgb=rgb, cwd=d), range(n_procs)))
        wall = time.perf_counter() - t0
    return n_procs * pics_per_proc / max(secs), wall


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per = args.ref_pics
    vals, walls = [], []
    for i in range(args.warmup + args.steps):
        fps, wall = reference_sample(cores, per)
        if i >= args.warmup:
            vals.append(fps); walls.append(wall)
    value = float(np.mean(vals))
    sample = f"{per} pictures per process x {cores} processes (one per host core), same 1080p stream, in-process minivideo_decode() time"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(walls)) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------
# our arm

def run_ours(args):
    rank, world, local = dist_env()
    if args.gpus > 1 and world == 1:
        # not launched by torchrun: re-exec under it (one process per GPU)
        port = 29500 + (os.getpid() % 1000)
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                   "--master-port", str(port), str(Path(__file__).resolve())] + sys.argv[1:])
    import torch
    import torch.distributed as dist
    from minivideo_b200 import api, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    try:        # run this rank (and allocate its pinned buffers) on the host cores next to its GPU
        import pynvml as nv
        nv.nvmlInit()
        nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation; stdout carries the one JSON line only
        sys.stdout.flush()
        keep = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(keep, 1)
            os.close(keep)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- synthetic input: G distinct pictures per rank, replicated on the device to F
    F, G, scale = args.frames, min(args.distinct, args.frames), args.rgb_scale
    _, soa = synth.generate(G, "1080p", want_stream=False, seed=0xC0FFEE + 2 + 1000 * rank)
    W, H, N = soa.width, soa.height, soa.n_mbs
    ctx = api.Context(local, soa.width_mbs, soa.height_mbs, F)
    ctx.set_sps_from(soa)
    ctx.upload(soa, 0)
    for s in range(G, F):
        ctx.clone_slot(s % G, s)
    ctx.sync()

    # ---- timed region: K steps over the resident batch
    sampler = ClockSampler(local)
    sampler.start()
    t_wait = time.perf_counter()
    while not sampler.samples and sampler.err is None and time.perf_counter() - t_wait < 10.0:
        time.sleep(0.01)                      # NVML initialisation can take longer than the warm-up
    split = args.pipeline == "split"
    ctx.set_pipeline_mode(api.PIPELINE_SPLIT if split else api.PIPELINE_FUSED)
    step = (lambda: ctx.run(0, F, scale)) if (split or scale != 1) else (lambda: ctx.run_rgb(0, F))
    for _ in range(args.warmup):
        step()
    ctx.sync()
    barrier()
    k_ms = {"k1": [], "k2": [], "k3": [], "kf": []}
    launches = 0
    t_wall0 = time.perf_counter()
    ctx.mark(0)
    for _ in range(args.steps):
        step()
        t = ctx.timing()                      # waits for the step; per-kernel CUDA-event times
        k_ms["k1"].append(t.k1_dequant_idct_ms); k_ms["k2"].append(t.k2_wavefront_ms); k_ms["k3"].append(t.k3_rgb_ms)
        k_ms["kf"].append(t.fused_ms)
        launches += t.launches
    ctx.mark(1)
    dev_ms = ctx.mark_elapsed_ms()
    barrier()
    t_wall1 = time.perf_counter()
    wall_ms = (t_wall1 - t_wall0) * 1e3
    clocks = sampler.result(t_wall0, t_wall1)

    # ---- end to end: pinned host buffers -> H2D -> kernels -> D2H RGB in pinned host memory, through the
    # C ABI.  Headline: the packed transfer format (mvg_decode_host_packed, what the front end emits);
    # the dense SoA call (mvg_decode_host) is timed beside it.
    E = min(args.e2e_frames, F)
    rgb_px = (W // scale) * (H // scale) * 3
    rgb_out = api.PinnedArray((E, rgb_px), np.uint8)
    d2h = rgb_out.nbytes
    packed = api.Packed(soa, n_pics=E, pinned=True)
    h2d = packed.nbytes
    for _ in range(max(1, args.warmup // 2)):
        ctx.decode_host_packed(packed, None, rgb_out.array, scale)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.decode_host_packed(packed, None, rgb_out.array, scale)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    reps = -(-E // G)
    pin = {
        "mb_kind": api.PinnedArray((E * N,), np.uint8), "i16_mode": api.PinnedArray((E * N,), np.uint8),
        "chroma_mode": api.PinnedArray((E * N,), np.uint8), "qp_y": api.PinnedArray((E * N,), np.int8),
        "cbp": api.PinnedArray((E * N,), np.uint8), "luma_modes": api.PinnedArray((E * N, 16), np.uint8),
        "coeff": api.PinnedArray((E * N, 384), np.int16),
    }
    for name, pa in pin.items():
        src = getattr(soa, name)
        pa.array[...] = np.concatenate([src] * reps)[: E * N]
    batch = api.Batch()
    batch.n_pics = E
    for name, pa in pin.items():
        setattr(batch, name, pa.ptr)
    h2d_dense = sum(pa.nbytes for pa in pin.values())
    ctx.decode_host(None, None, rgb_out.array, scale, batch=batch)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.decode_host(None, None, rgb_out.array, scale, batch=batch)
    torch.cuda.synchronize()
    e2e_dense_s = time.perf_counter() - t0

    # ---- from the bitstream: Annex-B bytes -> host front end (CAVLC on all host cores, packed output into pinned
    # memory) -> mvg_decode_host_packed -> RGB24 in pinned host memory.  Same input and output as the reference
    # arm (which times minivideo_decode() on the same kind of stream); bound by the host's entropy decoding.
    stream_e2e = None
    if rank == 0 and args.stream_frames > 0:
        import ctypes as C
        from minivideo_b200 import front
        S = min(args.stream_frames, F)
        stream, _ = synth.generate(S, "1080p", want_soa=False, seed=0xC0FFEE + 2)
        st = front.Stream(stream)
        ls4, ls8 = st.level_scale()
        ctx.set_sps(st.info.width_mbs, st.info.height_mbs, ls4, ls8, st.info.cb_qp_offset, st.info.cr_qp_offset)
        cap = S * N * 160                                       # words; ~5x what this stream needs
        order = ("mb_kind", "i16_mode", "chroma_mode", "qp_y", "luma_modes", "nz_blocks", "word_off", "pic_off", "words")

        def buffers():
            pk = {"mb_kind": api.PinnedArray((S * N,), np.uint8), "i16_mode": api.PinnedArray((S * N,), np.uint8),
                  "chroma_mode": api.PinnedArray((S * N,), np.uint8), "qp_y": api.PinnedArray((S * N,), np.int8),
                  "luma_modes": api.PinnedArray((S * N, 16), np.uint8), "nz_blocks": api.PinnedArray((S * N,), np.uint32),
                  "word_off": api.PinnedArray((S * N,), np.uint32), "pic_off": api.PinnedArray((S + 1,), np.uint64),
                  "words": api.PinnedArray((cap,), np.uint16)}
            return pk, front.FrontPackedBatch(S, *(pk[k].ptr for k in order), cap), api.PackedBatch(S, *(pk[k].ptr for k in order))
        sets = [buffers(), buffers()]                           # parse of step k+1 runs while step k is on the GPU
        out_s = api.PinnedArray((S, rgb_px), np.uint8)
        threads = os.cpu_count() or 1
        flib = front.lib()

        def parse(k):
            rc = flib.mvf_parse_pictures_packed(st.handle, None, 0, S, C.byref(sets[k & 1][1]), threads)
            if rc != 1:
                raise RuntimeError(flib.mvf_last_error(st.handle).decode())

        def decode(k):
            ctx._ck(ctx.lib.mvg_decode_host_packed(ctx.handle, C.byref(sets[k & 1][2]), None, out_s.ptr, scale))
        parse(0); decode(0)                                     # warm-up
        with ThreadPoolExecutor(max_workers=1) as ex:           # ctypes calls release the GIL
            t0 = time.perf_counter()
            parse(0)
            for k in range(args.steps):
                nxt = ex.submit(parse, k + 1) if k + 1 < args.steps else None
                decode(k)
                if nxt:
                    nxt.result()
            dt = time.perf_counter() - t0
        stream_e2e = {"value": S * args.steps / dt, "unit": UNIT, "pictures_per_step": S, "host_threads": threads,
                      "stream_bytes_per_picture": len(stream) // S,
                      "path": "Annex-B bytes -> mvf_parse_pictures_packed (CAVLC on the host cores) -> mvg_decode_host_packed -> RGB24; "
                              "the parse of step k+1 overlaps the GPU work of step k; one GPU, rank 0 only"}

    # ---- reduce over ranks (max time), rank 0 reports
    times = torch.tensor([dev_ms, wall_ms, e2e_s * 1e3, e2e_dense_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, e2e_ms, e2e_dense_ms = (float(x) for x in times.tolist())

    if rank == 0:
        peak, peak_src = measured_peak()
        ab = algorithmic_bytes(N, W, H, scale)
        mean_ms = {k: float(np.mean(v)) for k, v in k_ms.items()}
        ab = {k: v for k, v in ab.items() if mean_ms[k] > 0}
        mean_ms = {k: v for k, v in mean_ms.items() if v > 0}
        dom = max(mean_ms, key=mean_ms.get)
        achieved = ab[dom] * F / (mean_ms[dom] * 1e-3) / 1e9
        traffic = None
        tfile = ROOT / "profiles" / "ncu_traffic.json"
        if tfile.exists():
            try:
                traffic = json.loads(tfile.read_text())["per_picture_dram_bytes"][dom] * F
            except Exception:
                traffic = None
        kernels = {k: {"ms_per_launch": mean_ms[k], "algorithmic_bytes": ab[k] * F,
                       "achieved_gbs": ab[k] * F / (mean_ms[k] * 1e-3) / 1e9 if mean_ms[k] > 0 else None,
                       "frac": ab[k] * F / (mean_ms[k] * 1e-3) / 1e9 / peak if mean_ms[k] > 0 else None}
                   for k in mean_ms}
        line = {
            "metric": METRIC, "value": world * F * args.steps / (dev_ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pictures_per_step_per_gpu": F, "distinct_pictures": G,
                       "coded_size": f"{W}x{H}", "rgb_scale": scale, "pipeline": "3 kernels (k1 dequant/idct, k2 wavefront -> macroblock tiles, k3 rgb)",
                       "l2": f"inputs {F * N * 789 / 1e9:.1f} GB per step >> 126 MB L2 (no flush needed)",
                       "timing": "CUDA events on the launch stream, max over ranks", "wall_ms_per_step": wall_ms / args.steps},
            "clocks": clocks,
            "e2e": {"value": world * E * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "pictures_per_step_per_gpu": E,
                    "path": "mvg_decode_host_packed: pinned host packed SoA (sparse levels) -> H2D -> k0 expand, k1, k2, k3 -> D2H RGB24",
                    "dense": {"value": world * E * args.steps / (e2e_dense_ms * 1e-3), "h2d_bytes_per_step": h2d_dense,
                              "path": "mvg_decode_host: dense int16[384] levels per macroblock"}},
            "stream_e2e": stream_e2e,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": {"k1": "k1_dequant_idct", "k2": "k2_wavefront", "k3": "k3_rgb", "kf": "kf_recon"}[dom],
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src},
            "kernels": kernels,
            # the three stages together: algorithmic bytes of all kernels over the sum of their launch times
            "roofline_all_stages": {"achieved": sum(ab.values()) * F / (sum(mean_ms.values()) * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": sum(ab.values()) * F / (sum(mean_ms.values()) * 1e-3) / 1e9 / peak},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cores = os.cpu_count() or 1
                fps, _ = reference_sample(cores, args.ref_pics)
                line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sample": f"{args.ref_pics} pictures per process x {cores} processes of the same 1080p "
                                                  "stream through the unmodified reference (parse + recon + mb_to_rgb)"}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1000, help="pictures per step per GPU (resident in HBM)")
    ap.add_argument("--distinct", type=int, default=32, help="distinct pictures generated on the host per GPU")
    ap.add_argument("--rgb-scale", type=int, default=1, help="RGB thumbnail downscale factor (1 = the reference's mb_to_rgb)")
    ap.add_argument("--e2e-frames", type=int, default=384, help="pictures per end-to-end step per GPU")
    ap.add_argument("--stream-frames", type=int, default=64, help="pictures of the bitstream-to-RGB measurement (0 = skip)")
    ap.add_argument("--ref-pics", type=int, default=6, help="pictures each host core decodes in the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline", default="fused", choices=["fused", "split"], help="fused kernel (default) or round 1's three kernels")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
