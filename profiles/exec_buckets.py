"""Static SASS instructions of one kernel grouped by how often a warp executed them (ncu source-page CSV, see
summarize_source.py): separates the stages of a software-pipelined kernel, whose loops run different numbers of times.
python exec_buckets.py f.csv lo:hi:name ..."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
buckets = []
for a in sys.argv[2:]:
    lo, hi, name = a.split(":")
    buckets.append((int(lo), int(hi), name))
hdr = None


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


seen = set()
insts = []
for r in rows:
    if len(r) > 8 and r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] == '-' or not r[2].startswith("0x"):
        continue
    if r[2] in seen:
        continue
    seen.add(r[2])
    insts.append((int(r[2], 16), I(r[hdr.index('Instructions Executed')]), I(r[hdr.index('Predicated-On Thread Instructions Executed')]), r[3].strip()))
insts.sort()
tot = sum(i[1] for i in insts)
print(len(insts), "static instructions,", tot, "executed (warp level)")
out = {}
for a, c, t, s in insts:
    k = next((n for lo, hi, n in buckets if lo <= c < hi), "other")
    v = out.setdefault(k, [0, 0, 0])
    v[0] += c
    v[1] += 1
    v[2] += t
for k, v in out.items():
    print(f"{k:24s} {v[0]:>12d} executed ({100 * v[0] / tot:4.1f}%)  {v[1]:5d} static  avg lanes {v[2] / max(v[0], 1):.1f}")
