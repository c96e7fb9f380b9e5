#!/bin/bash
# ncu --set full of the hand-written FFT passes: one fft3_x2z + one ifft3_z2x at n^3 after a warm-up pair (6 kernels).
# Usage: bash tools/ncu_fft.sh <tag> [n]
tag=$1; n=${2:-512}
mkdir -p gpurun_out
cat > /tmp/ncu_fft_driver.py <<PY
import sys, torch
sys.path.insert(0, ".")
import padeops_b200 as pdo
pdo.decomp_2d.comm_init()
n = $n
ft = pdo.fft_3d(); assert ft.init(n, n, n, "x", 0.1, 0.1, 0.1) == 0
f = torch.randn(n, n, n, device="cuda", dtype=torch.float64)
for _ in range(2):
    s = ft.fft3_x2z(f); b = ft.ifft3_z2x(s)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:fft_ -s 6 -c 6 -f -o gpurun_out/${tag} python /tmp/ncu_fft_driver.py > gpurun_out/${tag}.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
tail -2 gpurun_out/${tag}.log
