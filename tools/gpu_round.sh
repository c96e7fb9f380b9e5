#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list and one --set full capture.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_round.sh <tag>'
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${tag}_gpu.txt
( time timeout 480 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
cat gpurun_out/${tag}_bench_n1.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches_bench_n1024.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-substep --profile-region \
    > gpurun_out/${tag}_launches.log 2>&1
timeout 300 bash tools/ncu_top.sh ${tag}
for sch in 1 2; do timeout 120 python tools/substep_bench.py 256 $sch 5; done > gpurun_out/${tag}_substep_n1.jsonl 2>&1
timeout 120 python tools/substep_bench.py 512 2 3 >> gpurun_out/${tag}_substep_n1.jsonl 2>&1
cat gpurun_out/${tag}_substep_n1.jsonl
ls -la gpurun_out | head -30
