"""BASELINE.json configs[4]: a 3840x2160 High-profile intra stream, 'distributed' thumbnail extraction, all GPUs of the
box (`mv_thumbnailer -d all`) beside the reference CLI on the same stream.  Development aid; prints wall-clock times.
    python tests/tools/config4_demo.py [n_pictures_in_stream] [n_extracted]"""
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from minivideo_b200 import synth  # noqa: E402

n, want = (int(sys.argv[1]) if len(sys.argv) > 1 else 48), (int(sys.argv[2]) if len(sys.argv) > 2 else 24)
t = time.time()
stream, _ = synth.generate(n, "2160p", seed=404)
print(f"stream: {n} pictures 3840x2160, {len(stream) / 1e6:.1f} MB, generated in {time.time() - t:.1f} s", flush=True)
with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
    src = Path(d) / "in.264"
    src.write_bytes(stream)
    outs = {}
    for name, exe, extra in (("reference", ROOT / "oracle" / "_ref" / "mini_thumbnailer", []),
                             ("ours, 1 GPU", ROOT / "minivideo_b200" / "mv_thumbnailer", ["-o", ".", "-d", "0"]),
                             ("ours, all GPUs", ROOT / "minivideo_b200" / "mv_thumbnailer", ["-o", ".", "-d", "all", "-b", "4"])):
        out = Path(d) / "out"
        out.mkdir()
        t = time.time()
        r = subprocess.run([str(exe), "-i", str(src), "-f", "yuv420", "-n", str(want), "-e", "distributed"] + extra, cwd=out,
                           capture_output=True, text=True, env=dict(os.environ, MVT_TIMING="1"))
        dt = time.time() - t
        files = {p.name: p.stat().st_size for p in out.iterdir() if not p.name.startswith("core")}
        full = {k for k, v in files.items() if v == 3840 * 2160 * 3 // 2}
        outs[name] = {k: (out / k).read_bytes() for k in full}
        print(f"{name:15s}: rc {r.returncode}, {len(full)} complete files in {dt:6.2f} s = {len(full) / dt:6.1f} pictures/s", flush=True)
        for line in r.stderr.splitlines():
            if line.startswith("mvt_extract["):
                print("    ", line)
        subprocess.run(["rm", "-rf", str(out)])
    ref = outs["reference"]
    for name in ("ours, 1 GPU", "ours, all GPUs"):
        same = [k for k in ref if outs[name].get(k) == ref[k]]
        print(f"{name}: {len(same)} of {len(ref)} files of the reference are byte-identical")
