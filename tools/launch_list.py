#!/usr/bin/env python
"""Print the launches of the LAST iteration in an `ncu --metrics gpu__time_duration.sum --csv` log (from the last launch of
`--first` on): python tools/launch_list.py gpurun_out/c3_launches.csv --first volume_kernel"""
import argparse
import csv

ap = argparse.ArgumentParser()
ap.add_argument("path")
ap.add_argument("--first", default="volume_kernel")
ap.add_argument("--width", type=int, default=100)
a = ap.parse_args()
rows = list(csv.reader(open(a.path)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
L = [(r[kn], float(r[mv].replace(",", "")) / 1e3) for r in rows[hi + 2:] if len(r) == len(hdr)]
idx = [i for i, (k, _) in enumerate(L) if a.first in k]
last = L[idx[-1]:] if idx else L
try:
    print("launches %d, total %.1f us" % (len(last), sum(t for _, t in last)))
    for k, t in last:
        print("%8.1f  %s" % (t, k[:a.width]))
except BrokenPipeError:                                  # piped into head
    pass
