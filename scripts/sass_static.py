"""Dev helper: static SASS instruction count of one kernel, by source line bucket.
    python scripts/sass_static.py <kernel-name-substring> [bucket]   (reads minivideo_b200/libmvgpu.so)"""
import collections, re, subprocess, sys, tempfile, os
name = sys.argv[1]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 20
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath("minivideo_b200/libmvgpu.so")], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cub], cwd=d, capture_output=True, text=True).stdout
cur, on, cnt, ops = None, False, collections.Counter(), collections.Counter()
for l in txt.splitlines():
    if l.startswith(".text."):
        on = name in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        cnt[cur] += 1
        ops[m.group(2).split(".")[0]] += 1
print("total", sum(cnt.values()), "instructions =", sum(cnt.values()) * 16 // 1024, "KB")
b = collections.Counter()
for (f, n), c in cnt.items():
    b[(f, n // step * step)] += c
for k in sorted(b):
    print(f"{k[0]:20s} {k[1]:5d} {b[k]:5d}")
print(dict(ops.most_common(25)))
