#!/usr/bin/env python
"""Host-CPU probe for the reference arm: usable cores (affinity / cgroup quota) and how the oracle GwcNet forward
scales with the torch thread count on this box."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "usable", bench.usable_cores())
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    if os.path.exists(p):
        print(p, open(p).read().strip())
print("loadavg", open("/proc/loadavg").read().strip())
model = bench.oracle_model()
x = {"left": torch.randn(1, 3, 256, 512), "right": torch.randn(1, 3, 256, 512)}
for n in [int(a) for a in sys.argv[1:]] or [8, 16, 32, 64]:
    torch.set_num_threads(n)
    with torch.no_grad():
        model(dict(x))
        t0 = time.perf_counter()
        model(dict(x))
        print("threads", n, "s/pair", round(time.perf_counter() - t0, 3), flush=True)
