"""Static SASS evidence for profiles/: per kernel of libmvgpu.so the instruction count, registers and the opcode histogram
(cuobjdump -sass), with the mnemonics that show what the kernel is built from (UBLKCP = 1-D TMA bulk copy, SYNCS =
mbarrier, LDGSTS = per-lane async copy, VIADDMNMX = DPX clip-add, IDP = dp4a, no HMMA/UTCMMA: nothing is a contraction).
    python scripts/sass_hist.py > profiles/r02_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "minivideo_b200", "libmvgpu.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
regs = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", res))
kern, ops = None, collections.defaultdict(collections.Counter)
for l in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        kern = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m and kern:
        ops[kern][m.group(1)] += 1
print("arch:", ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", txt)))))
for k, c in ops.items():
    n = sum(c.values())
    name = subprocess.run(["cu++filt", k], capture_output=True, text=True).stdout.strip() or k
    print(f"\n{name}\n  {n} instructions ({n * 16 // 1024} KB), {regs.get(k, '?')} registers")
    print("  " + ", ".join(f"{o} {v}" for o, v in c.most_common()))
    marks = {o: c[o] for o in ("UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "VIADDMNMX", "VIMNMX", "IDP", "PRMT", "SHFL", "VOTE", "REDUX", "NANOSLEEP", "HMMA", "UTCMMA", "IMMA") if c[o]}
    print("  markers:", marks)
