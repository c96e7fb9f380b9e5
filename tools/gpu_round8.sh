#!/bin/bash
# One 8-GPU gpurun call: multi-rank parity (tests/test_multigpu.py at 2 / 4 / 8 ranks, worker logs kept), the three NVLink data
# planes of the fused transposes at 1024^3 on 1x8 and 2x4, and the bench line at N = 8 with the best plane.
# Usage: gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_round8.sh <tag>'
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo8.txt 2>&1
( time timeout 700 python -m pytest tests/test_multigpu.py -m gpu -q -rf -p no:cacheprovider ) > gpurun_out/${tag}_pytest_mgpu8.log 2>&1
tail -6 gpurun_out/${tag}_pytest_mgpu8.log | cut -c1-500
: > gpurun_out/${tag}_transposes_8gpu.jsonl
for plane in ce bulk sm; do
  PDO_P2P_MODE=$plane timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 \
      tools/transpose_bench.py 1024 --grids=1x8,2x4 2>&1 | grep "^{" | sed "s/\"path\": \"p2p\"/\"path\": \"p2p-$plane\"/" >> gpurun_out/${tag}_transposes_8gpu.jsonl
done
python - ${tag} <<'PY' > gpurun_out/${tag}_best_plane.txt
import json, sys, collections
worst = collections.defaultdict(lambda: 1.0)
for l in open("gpurun_out/%s_transposes_8gpu.jsonl" % sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_transposes_8gpu.jsonl"):
    d = json.loads(l)
    if d["nvlink_frac"] > 0:
        worst[d["path"]] = min(worst[d["path"]], d["nvlink_frac"])
best = max(worst, key=worst.get) if worst else "p2p-ce"
print(best.split("-")[1])
for k, v in sorted(worst.items()):
    print(k, v, file=sys.stderr)
PY
plane=$(cat gpurun_out/${tag}_best_plane.txt)
echo "best plane: $plane"
cut -c1-220 gpurun_out/${tag}_transposes_8gpu.jsonl | awk 'NR<=48'
PDO_P2P_MODE=$plane timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_bench_n8.err
tail -3 gpurun_out/${tag}_bench_n8.err | cut -c1-300
cut -c1-7000 gpurun_out/${tag}_bench_n8.json
