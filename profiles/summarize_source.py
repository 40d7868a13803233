"""Per-source-line totals of an ncu source page: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name K > f.csv ; python summarize_source.py f.csv [top]"""
import csv
import sys


def main(path, top=45):
    rows = list(csv.reader(open(path, errors="replace")))
    files, cur = {}, None
    hdr = None
    per = {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[2] != "-":      # only the per-source-line summary rows (Address == "-")
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        inst = int(r[hdr.index("Instructions Executed")] or 0)
        samp = int(r[hdr.index("# Samples")] or 0)
        k = (cur, line)
        a = per.setdefault(k, [0, 0, r[1].strip()[:110]])
        a[0] += inst
        a[1] += samp
    tot_i = sum(v[0] for v in per.values()) or 1
    tot_s = sum(v[1] for v in per.values()) or 1
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    for (f, line), (i, s, src) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*i/tot_i:5.1f}% inst {100*s/tot_s:5.1f}% smp  {f}:{line:<5d} {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
