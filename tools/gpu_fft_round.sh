#!/bin/bash
# One gpurun call for the hand-written FFT passes: parity tests, pass timings against cuFFT, the substep both ways, launch list.
tag=${1:-r02p}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_spectral_gpu.py tests/test_igrid_gpu.py -m gpu -x -q ) > gpurun_out/${tag}_pytest_fft.log 2>&1
tail -15 gpurun_out/${tag}_pytest_fft.log
timeout 200 python tools/fftbench.py 256 512 > gpurun_out/${tag}_fftbench.jsonl 2>&1
PDO_FFT=cufft timeout 200 python tools/fftbench.py 256 512 >> gpurun_out/${tag}_fftbench.jsonl 2>&1
cat gpurun_out/${tag}_fftbench.jsonl
for sch in 1 2; do timeout 150 python tools/substep_bench.py 512 $sch 3; done > gpurun_out/${tag}_substep_n1.jsonl 2>&1
PDO_FFT=cufft timeout 150 python tools/substep_bench.py 512 1 3 >> gpurun_out/${tag}_substep_n1.jsonl 2>&1
timeout 150 python tools/substep_bench.py 256 1 5 >> gpurun_out/${tag}_substep_n1.jsonl 2>&1
cat gpurun_out/${tag}_substep_n1.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_substep_n512.csv \
    python tools/substep_bench.py 512 1 2 > gpurun_out/${tag}_launches_substep.log 2>&1
tail -2 gpurun_out/${tag}_launches_substep.log
