"""Wall-clock of the mv_thumbnailer CLI (bitstream file -> picture files on /dev/shm) against the reference CLI
on the same 1080p stream.  Development aid; numbers go to profiles/ by hand.
    python tests/tools/cli_timing.py [n_pictures] [n_reference_pictures]"""
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from minivideo_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_ref = int(sys.argv[2]) if len(sys.argv) > 2 else 8
t = time.time()
stream, _ = synth.generate(n, "1080p", profile_idc=100, transform8x8=1, scaling_lists=1, seed=77)
print(f"stream: {n} pictures, {len(stream) / 1e6:.1f} MB, generated in {time.time() - t:.1f} s", flush=True)
with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
    src = Path(d) / "in.264"
    src.write_bytes(stream)
    batches = [b for b in os.environ.get("MVT_BATCHES", "64").split(",")]
    for fmt in ("yuv420", "bmp", "png"):
        for exe, cnt, b in [(ROOT / "minivideo_b200" / "mv_thumbnailer", n, b) for b in batches] + [(ROOT / "oracle" / "_ref" / "mini_thumbnailer", n_ref, None)]:
            if not exe.exists():
                continue
            out = Path(d) / "out"
            out.mkdir()
            extra = ["-o", ".", "-b", b] if exe.name == "mv_thumbnailer" else []
            t = time.time()
            r = subprocess.run([str(exe), "-i", str(src), "-f", fmt, "-n", str(cnt)] + extra, cwd=out, capture_output=True, text=True,
                               env=dict(os.environ, MVT_TIMING="1"))
            dt = time.time() - t
            for line in r.stderr.splitlines():
                if line.startswith("mvt_extract["):
                    print("   ", line)
            files = [p for p in out.iterdir() if not p.name.startswith("core")]
            print(f"{exe.name:18s} -f {fmt:6s} -n {cnt:3d} -b {b}: rc {r.returncode} {len(files)} files in {dt:6.2f} s = {len(files) / dt:7.1f} pictures/s", flush=True)
            subprocess.run(["rm", "-rf", str(out)])
