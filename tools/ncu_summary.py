#!/usr/bin/env python
"""Summarise .ncu-rep files (read here, no GPU needed) into a markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/r1_volume.ncu-rep [more.ncu-rep ...] > profiles/r1_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        if len(r) == len(hdr):
            yield hdr, units, r


def main():
    for path in sys.argv[1:]:
        print("## %s\n" % path.split("/")[-1])
        for hdr, units, r in rows_of(path):
            name = r[hdr.index("Kernel Name")]
            print("**%s**\n" % name)
            print("| metric | value |\n|---|---|")
            for key, label in METRICS:
                if key in hdr:
                    i = hdr.index(key)
                    print("| %s (`%s`) | %s %s |" % (label, key, r[i], units[i]))
            print()


if __name__ == "__main__":
    main()
