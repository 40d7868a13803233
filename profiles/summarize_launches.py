"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals/shares and one steady-state frame."""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        seq.append((row["Kernel Name"].split("(")[0], v))
    return seq


def main(path):
    seq = load(path)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in seq:
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(seq)} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} n={v[0]:4d} total={v[1]:10.1f} us avg={v[1] / v[0]:9.2f} us share={100 * v[1] / tot:5.1f}%")
    idx = [i for i, (n, _) in enumerate(seq) if n.endswith("k_shade")]
    if len(idx) >= 3:
        i0, i1 = idx[-3] + 1, idx[-2] + 1
        ft = sum(v for _, v in seq[i0:i1])
        print(f"# one steady-state frame ({i1 - i0} launches, {ft:.1f} us):")
        for n, v in seq[i0:i1]:
            print(f"   {n:40s} {v:9.2f} us {100 * v / ft:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
