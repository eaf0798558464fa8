"""Dev helper: top CUDA source lines by executed instructions for one kernel of an .ncu-rep."""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[2]
idx = {}
for i, h in enumerate(hdr):
    idx.setdefault(h, i)
iI, iS = idx['Instructions Executed'], idx['# Samples']
def f(x):
    try: return int(float(x))
    except Exception: return 0
lines = [(int(r[0]), r[1].strip()[:115], f(r[iI]), f(r[iS])) for r in rows[3:] if len(r) > iI and r[0].isdigit()]
tot, ts = sum(l[2] for l in lines), sum(l[3] for l in lines)
print('total inst', tot, 'samples', ts)
for ln, src, inst, samp in sorted(lines, key=lambda l: -l[2])[:top]:
    print(f'{ln:5d} {100*inst/tot:5.1f}% inst {100*samp/max(ts,1):5.1f}% samp | {src}')
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
sass = [r for r in rows[3:] if len(r) > iI and r[0] == '']
agg = {h: sum(f(r[idx[h]]) for r in sass) for h in st}
tt = sum(agg.values())
print({k: round(100*v/tt, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
from collections import Counter
c = Counter()
for r in sass:
    toks = r[3].split() if len(r) > 3 else []
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    c[op.split('.')[0]] += f(r[iI])
print({k: round(100*v/tot, 1) for k, v in c.most_common(18)})
