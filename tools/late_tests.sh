#!/bin/bash
# First hardware run of the GPU tests written after a round's last GPU session: they carry xfail(strict=False), so the ordinary
# `pytest -m gpu` run can never turn red on them; here they run with --runxfail (real failures, full tracebacks) and the log goes
# to gpurun_out/.  Once green, remove their xfail marks.
# Usage: gpurun --timeout 900 -- 'bash tools/late_tests.sh <tag>'        (add --gpus 2 for the multi-GPU late sections)
tag=${1:-late}
mkdir -p gpurun_out
files="tests/test_ops_periodic_gpu.py tests/test_stagg_nonperiodic_gpu.py tests/test_igrid_gpu.py tests/test_vecops_gpu.py \
       tests/test_spectral_gpu.py tests/test_nonperiodic_gpu.py tests/test_multigpu.py"
( time timeout 800 python -m pytest $files -m gpu --runxfail -q -rf ) > gpurun_out/${tag}_late_tests.log 2>&1
tail -40 gpurun_out/${tag}_late_tests.log
