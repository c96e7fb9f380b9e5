#!/bin/bash
# ncu --set full of ONE operator call with a forced kernel variant.
# Usage: bash tools/ncu_one.sh <tag> <strided_mode|-> <x_threads|-> <op> <axis> <n> <na> <nb>
tag=$1; mode=$2; xth=$3; shift 3
[ "$mode" != "-" ] && export PDO_STRIDED_MODE=$mode
[ "$xth" != "-" ] && export PDO_X_THREADS=$xth
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:chunk -s 4 -c 1 -f -o gpurun_out/${tag} python tools/profile_one.py "$@" > gpurun_out/${tag}.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
tail -2 gpurun_out/${tag}.log
