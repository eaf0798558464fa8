"""Robustness of the host front end: corrupted streams (bit flips, truncation, garbage runs) under
AddressSanitizer + UBSan must end in a return code, never in a crash or an out-of-bounds access."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("cfg", [dict(width_mbs=6, height_mbs=4, profile_idc=66, seed=41),
                                 dict(width_mbs=5, height_mbs=5, profile_idc=100, transform8x8=1, scaling_lists=1, seed=42)])
def test_front_end_survives_corrupted_streams_under_sanitizers(tmp_path, cfg):
    from minivideo_b200 import synth
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = tmp_path / "fuzz_front"
    r = subprocess.run(["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                        f"-I{ROOT / 'include'}", f"-I{ROOT / 'minivideo_b200' / 'csrc'}", "-o", str(exe),
                        str(ROOT / "tests" / "fuzz_front.c"), str(ROOT / "minivideo_b200" / "csrc" / "h264_front.c"),
                        "-lm", "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build not available: " + r.stderr[-300:])
    stream, _ = synth.generate(4, **cfg)
    (tmp_path / "s.264").write_bytes(stream)
    r = subprocess.run([str(exe), str(tmp_path / "s.264"), "1500"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    assert "rounds=1500" in r.stdout
