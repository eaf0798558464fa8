"""TEST INFRASTRUCTURE: run the compiled, unmodified reference decoder (oracle/_ref,
built by oracle/Makefile) on an Annex-B stream and read back what it produced."""
from __future__ import annotations

import os
import re
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DECODE = HERE / "_ref" / "ref_decode"
MINI_THUMBNAILER = HERE / "_ref" / "mini_thumbnailer"


def available() -> bool:
    return REF_DECODE.exists() and os.access(REF_DECODE, os.X_OK)


def _tmpdir():
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    return tempfile.TemporaryDirectory(dir=base, prefix="mvref_")


def decode(stream: bytes, n_pics: int, width: int, height: int, want_rgb: bool = False,
           want_soa: bool = False):
    """Decode with ref_decode (reference objects + export_idr tap).

    Returns dict(yuv=[P, 1.5*W*H] u8, rgb=[P, H, W, 3] u8 | None, soa=raw bytes | None)."""
    with _tmpdir() as d:
        src = Path(d) / "in.264"
        src.write_bytes(stream)
        cmd = [str(REF_DECODE), str(src), str(n_pics), "--yuv", str(Path(d) / "o.yuv")]
        if want_rgb:
            cmd += ["--rgb", str(Path(d) / "o.rgb")]
        if want_soa:
            cmd += ["--soa", str(Path(d) / "o.soa")]
        res = subprocess.run(cmd, capture_output=True, text=True, cwd=d)
        if res.returncode != 0:
            raise RuntimeError(f"ref_decode failed rc={res.returncode}: {res.stdout[-2000:]} {res.stderr[-2000:]}")
        yuv = np.fromfile(Path(d) / "o.yuv", np.uint8).reshape(n_pics, width * height * 3 // 2)
        rgb = np.fromfile(Path(d) / "o.rgb", np.uint8).reshape(n_pics, height, width, 3) if want_rgb else None
        soa = (Path(d) / "o.soa").read_bytes() if want_soa else None
    return dict(yuv=yuv, rgb=rgb, soa=soa)


def decode_cli(stream: bytes, n_pics: int, width: int, height: int) -> np.ndarray:
    """Decode through the reference's own CLI (`mini_thumbnailer -f yuv420`), i.e.
    export_idr_yuv420() writing files into the CWD (export.c:627-642)."""
    with _tmpdir() as d:
        src = Path(d) / "in.264"
        src.write_bytes(stream)
        res = subprocess.run([str(MINI_THUMBNAILER), "-i", str(src), "-f", "yuv420", "-n", str(n_pics),
                              "-e", "unfiltered"], capture_output=True, text=True, cwd=d)
        if res.returncode != 0:
            raise RuntimeError(f"mini_thumbnailer failed: {res.stdout[-2000:]} {res.stderr[-2000:]}")
        out = []
        for i in range(n_pics):
            name = "in.yuv" if n_pics == 1 else f"in_{i}.yuv"
            out.append(np.fromfile(Path(d) / name, np.uint8))
        return np.stack(out).reshape(n_pics, width * height * 3 // 2)


def time_decode(stream_path: str, n_pics: int, rgb: bool = True, cwd: str | None = None) -> float:
    """Wall seconds the reference spends in minivideo_decode() for n_pics pictures
    (CAVLC parse + reconstruction + mb_to_rgb when rgb=True; no file output)."""
    cmd = [str(REF_DECODE), stream_path, str(n_pics), "--time"] + ([] if rgb else ["--norgb"])
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=cwd)
    m = re.search(r"REFTIME pictures=(\d+) seconds=([0-9.]+)", res.stdout)
    if res.returncode != 0 or not m or int(m.group(1)) != n_pics:
        raise RuntimeError(f"ref_decode --time failed: {res.stdout[-1000:]} {res.stderr[-1000:]}")
    return float(m.group(2))


def parse_soa(raw: bytes):
    """Parse the --soa dump of ref_driver.c into a minivideo_b200.synth.Soa + tables."""
    from minivideo_b200.synth import Soa
    h = np.frombuffer(raw, np.int32, 6)
    assert h[0] == 0x4153564D
    w, hh, p, cb, cr = (int(x) for x in h[1:6])
    off = 24
    ls4 = np.frombuffer(raw, np.int32, 288, off).reshape(3, 6, 16); off += 288 * 4
    ls8 = np.frombuffer(raw, np.int32, 384, off).reshape(6, 64); off += 384 * 4
    n = w * hh
    fields = {k: [] for k in ("kind", "i16", "cm", "qp", "cbp", "modes", "coeff")}
    for _ in range(p):
        for k, cnt, dt in (("kind", n, np.uint8), ("i16", n, np.uint8), ("cm", n, np.uint8), ("qp", n, np.int8),
                           ("cbp", n, np.uint8), ("modes", n * 16, np.uint8), ("coeff", n * 384, np.int16)):
            a = np.frombuffer(raw, dt, cnt, off); off += a.nbytes
            fields[k].append(a)
    cat = {k: np.concatenate(v) for k, v in fields.items()}
    soa = Soa(w, hh, p, cat["kind"], cat["i16"], cat["cm"], cat["qp"], cat["cbp"],
              cat["modes"].reshape(-1, 16), cat["coeff"].reshape(-1, 384), cb_qp_offset=cb, cr_qp_offset=cr)
    return soa, ls4, ls8
