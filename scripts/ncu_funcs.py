"""Dev helper: attribute executed instructions of k2_wavefront to its device functions and list the
heaviest source lines of a line range."""
import csv, subprocess, io, re, sys, bisect
rep = sys.argv[1]; lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0; hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
nmb = float(sys.argv[4]) if len(sys.argv) > 4 else 384 * 8160
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "k2_wavefront"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[2]; idx = {}
for i, h in enumerate(hdr): idx.setdefault(h, i)
iI = idx['Instructions Executed']
def f(x):
    try: return int(float(x))
    except Exception: return 0
lines = [(int(r[0]), r[1].strip()[:118], f(r[iI])) for r in rows[3:] if len(r) > iI and r[0].isdigit()]
src = open('/root/repo/minivideo_b200/csrc/mvg_kernels.cuh').read().split('\n')
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r'^(?:__device__|__global__|template|k2_|static)?.*?\b(k2_[a-z0-9_]+)\(', l)
    if m and not l.startswith(' ') and not l.strip().startswith('//'): marks.append((i, m.group(1)))
starts = [m[0] for m in marks]
agg = {}
for ln, s_, inst in lines:
    k = bisect.bisect_right(starts, ln) - 1
    name = marks[k][1] if k >= 0 else 'helpers'
    agg[name] = agg.get(name, 0) + inst
tot = sum(l[2] for l in lines)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]): print(f'{k:20s} {100*v/tot:5.1f}%  {v/nmb:7.1f} instr/MB')
print('total/MB', tot / nmb)
for ln, s_, inst in sorted([l for l in lines if lo <= l[0] < hi], key=lambda l: -l[2])[:40]:
    print(f'{ln:5d} {inst/nmb:6.1f} | {s_}')
