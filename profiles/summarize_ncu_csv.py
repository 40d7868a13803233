"""Key metrics of an `ncu --set full` report already exported with `ncu -i X.ncu-rep --page raw --csv > X.csv` (what travels back
from the GPU box instead of the report): python summarize_ncu_csv.py X.csv [Y.csv ...]   — same block format as summarize_ncu.py"""
import csv
import sys

from summarize_ncu import WANT


def main(paths):
    for p in paths:
        rows = list(csv.reader(open(p, errors="replace")))
        rows = [r for r in rows if len(r) > 8]
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            print("=== " + r[hdr.index("Kernel Name")][:80])
            for w in WANT:
                if w in hdr:
                    print(f"  {w:86s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")


if __name__ == "__main__":
    main(sys.argv[1:])
