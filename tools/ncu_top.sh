#!/bin/bash
# One `ncu --set full` capture of the three CD10 kernels of the bench step (ddx, ddy, ddz at 1024^3) + raw CSV export.
# bench.py brackets its timed region with cudaProfilerStart/Stop, so exactly the timed launches are captured.
# Usage (under gpurun): bash tools/ncu_top.sh <tag>
set -u
tag=${1:-r01}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${tag}_cd10_n1024 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-substep --profile-region > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i gpurun_out/${tag}_cd10_n1024.ncu-rep --page raw --csv > gpurun_out/${tag}_cd10_n1024_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_cd10_n1024.ncu-rep --page details --csv > gpurun_out/${tag}_cd10_n1024_details.csv 2>/dev/null
