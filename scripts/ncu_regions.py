"""Dev helper: executed warp instructions of one profiled kernel, per enclosing source function (and per 8-line window
of the kernel body), normalised per macroblock.  Joins the SASS page of an .ncu-rep (executed counts, address order)
with `nvdisasm -g` of the SAME libmvgpu.so build (source lines, address order).
    python scripts/ncu_regions.py <rep> <kernel-substring> <n_macroblocks> [launch-index]
MVG_SO / MVG_SRC: the libmvgpu.so and the csrc directory of the profiled build (scripts/ncu_kf.sh keeps both)."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, name, n_mb = sys.argv[1], sys.argv[2], float(sys.argv[3])
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# ---- static: (file, line) per instruction
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("MVG_SO", os.path.join(ROOT, "minivideo_b200/libmvgpu.so"))], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cub], cwd=d, capture_output=True, text=True).stdout
static, on, cur = [], False, None
for l in txt.splitlines():
    if l.startswith(".text."):
        on = name in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f = m.group(1).split("/")[-1]
        if f.startswith("mvg_"):
            cur = (f, int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        static.append(cur)

# ---- dynamic: executed counts in address order (one launch)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + name.split("ILi")[0].lstrip("_Z0123456789")],
                     capture_output=True, text=True).stdout
blocks, cur_rows = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur_rows = []
        blocks.append([row[1], None, cur_rows])
    elif row and row[0] == "Address":
        blocks[-1][1] = row
    elif cur_rows is not None and row and row[0].startswith("0x"):
        cur_rows.append(row)
blocks = [b for b in blocks if len(b[2]) == len(static)]
if not blocks:
    sys.exit(f"no launch with {len(static)} instructions in the report (is the .so the profiled build?)")
kname, hdr, rows = blocks[min(which, len(blocks) - 1)]
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
iW, iX = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive")
print(kname, "launch", which, ":", len(rows), "SASS instructions")

# ---- source functions
funcs = {}
for f in ("mvg_kernels.cuh", "mvg_fused.cuh"):
    cur_f, table = "?", []
    for i, l in enumerate(open(os.path.join(os.environ.get("MVG_SRC", os.path.join(ROOT, "minivideo_b200/csrc")), f)), 1):
        m = re.match(r"^(?:template.*\n)?(?:__device__|__global__|static|template)[^;]*?\b([A-Za-z_0-9]+)\s*\(", l)
        if m and not l.startswith(" "):
            cur_f = m.group(1)
        m2 = re.match(r"^k[0-9f]_?[a-z_0-9]*\(|^([a-z_0-9]+)\(", l)
        if m2 and not l.startswith(" ") and "(" in l:
            cur_f = l.split("(")[0].split()[-1]
        table.append(cur_f)
    funcs[f] = table

by_f, by_w, samples, wave, wavex = (collections.Counter() for _ in range(5))
tot = tot_s = 0
for loc, r in zip(static, rows):
    n, s = int(float(r[iI] or 0)), int(float(r[iS] or 0))
    tot += n; tot_s += s
    fnw = funcs[loc[0]][loc[1] - 1] if loc else "?"
    wave[fnw] += int(float(r[iW] or 0)); wavex[fnw] += int(float(r[iX] or 0))
    fn = funcs[loc[0]][loc[1] - 1] if loc else "?"
    by_f[fn] += n; samples[fn] += s
    if loc:
        by_w[(loc[0], loc[1] // 8 * 8)] += n
print(f"total {tot / n_mb:7.1f} warp instructions per macroblock, {tot_s} samples")
for fn, n in by_f.most_common():
    print(f"  {fn:28s} {n / n_mb:7.1f}  {100 * n / tot:5.1f}% inst  {100 * samples[fn] / max(tot_s, 1):5.1f}% samples   "
          f"shared wavefronts {wave[fn] / n_mb:6.1f} (excess {wavex[fn] / n_mb:5.1f})")
print(f"  shared-memory wavefronts per macroblock: {sum(wave.values()) / n_mb:.1f}, of which bank conflicts {sum(wavex.values()) / n_mb:.1f}")
if True:
    pass
print("by 8-line window:")
for w in sorted(by_w):
    if by_w[w] / n_mb >= 1.0:
        print(f"  {w[0]:18s} line {w[1]:4d}  {by_w[w] / n_mb:7.1f}")
