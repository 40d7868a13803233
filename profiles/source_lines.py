"""SASS instructions (with warp- and thread-level execution counts) behind given source lines of an ncu source-page CSV
(see summarize_source.py): python source_lines.py f.csv 1264 1139 ..."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
want = set(sys.argv[2:])
hdr = None
cur = None


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


for r in rows:
    if len(r) > 8 and r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[2] == '-':
        cur = r[0]
        if cur in want:
            print("==", cur, r[1].strip()[:100])
        continue
    if cur in want:
        print("   %10d w %12d t  %s" % (I(r[hdr.index('Instructions Executed')]), I(r[hdr.index('Predicated-On Thread Instructions Executed')]), r[3].strip()[:70]))
