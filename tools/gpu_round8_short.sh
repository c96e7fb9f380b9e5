#!/bin/bash
# Trimmed 8-GPU call: multi-rank parity at 2 / 4 / 8 ranks, the bench line at N = 8 (with its substep / transposes / Poisson / 2048^3
# legs) and the igrid substep with and without the z-resident projection.
# Usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_round8_short.sh <tag>'
tag=${1:-r02z}
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q -rf -p no:cacheprovider ) > gpurun_out/${tag}_pytest_mgpu8.log 2>&1
tail -6 gpurun_out/${tag}_pytest_mgpu8.log | cut -c1-400
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_bench_n8.err
tail -3 gpurun_out/${tag}_bench_n8.err | cut -c1-300
cut -c1-6000 gpurun_out/${tag}_bench_n8.json
for z in 1 0; do
  PDO_IG_ZRESIDENT=$z timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2972$z \
      tools/substep_bench.py 512 1 3 2>&1 | grep "^{" | sed "s/^{/{\"zresident\": $z, /" | tee -a gpurun_out/${tag}_substep_n8.jsonl
done
