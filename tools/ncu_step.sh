#!/bin/bash
# One `ncu --set full` pass over EVERY kernel this library launches in one bench step (GwcNet, B = 8, 256x512, D = 192), run on the GPU box:
#     bash tools/ncu_step.sh <tag>          (e.g. r2_step)
# The report (too big for gpurun_out) stays in /tmp on the box; what comes back is
#     gpurun_out/<tag>_raw.csv            ncu --page raw --csv of every captured launch
#     gpurun_out/<tag>_summary.md         tools/ncu_summary.py over the same report
#     gpurun_out/<tag>_launches.csv       the cheap gpu__time_duration launch list of the same step (cuDNN kernels included)
#     gpurun_out/<tag>_conv3d_tc.ncu-rep  one launch of the dominant kernel with source correlation (-lineinfo build)
# Numbers printed by bench.py under ncu are never bench values.
set -u
TAG=${1:-step}
mkdir -p gpurun_out
export OSB_NCU_RANGE=1
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-comparators"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > /dev/null 2>&1
ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled -k regex:osb:: -f -o /tmp/${TAG} $BENCH > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py --aggregate /tmp/${TAG}.ncu-rep > gpurun_out/${TAG}_summary.md 2>/dev/null
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:conv3d_tc_kernel -c 1 -f \
    -o gpurun_out/${TAG}_conv3d_tc $BENCH > /dev/null 2>&1
ls -la gpurun_out/${TAG}_* /tmp/${TAG}.ncu-rep
