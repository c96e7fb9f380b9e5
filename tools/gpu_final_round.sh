#!/bin/bash
# End-of-round single-GPU call: the whole -m gpu suite, the bench line, launch lists (bench region, igrid substep), one --set full
# capture of the FFT passes.
tag=${1:-r02f}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -4 gpurun_out/${tag}_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
cut -c1-400 gpurun_out/${tag}_bench_n1.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_reference_arm_n1.json 2>/dev/null
cut -c1-300 gpurun_out/${tag}_reference_arm_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_substep_n512.csv \
    python tools/substep_bench.py 512 1 2 > gpurun_out/${tag}_launches_substep.log 2>&1
tail -1 gpurun_out/${tag}_launches_substep.log
timeout 150 python tools/substep_bench.py 512 1 3 hit
