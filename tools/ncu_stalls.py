#!/usr/bin/env python
"""Top stall locations of a source-correlated .ncu-rep (read here, no GPU): samples per CUDA source line with the leading reasons.

    python tools/ncu_stalls.py gpurun_out/r2_conv6.ncu-rep [top=30]
"""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    hdr, cur, data = None, None, []
    for r in csv.reader(io.StringIO(out)):
        if len(r) >= 2 and r[0] in ("File Path", "File Name"):
            cur = r[1].split("/")[-1]
        elif len(r) > 6 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0] != "":
            try:
                data.append((int(r[hdr.index("# Samples")]), cur, r))
            except ValueError:
                pass
    tot = sum(n for n, _, _ in data) or 1
    cols = [j for j, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
    print("%s: %d samples" % (path, tot))
    for n, f, r in sorted(data, key=lambda x: -x[0])[:top]:
        st = sorted([(int(r[j] or 0), hdr[j][6:]) for j in cols], reverse=True)[:3]
        print("%6d %5.1f%% %-14s:%-4s %-100s %s" % (n, 100.0 * n / tot, f, r[0], r[1].strip()[:100], [(a, b) for a, b in st if a]))


if __name__ == "__main__":
    main()
