#!/bin/bash
# round 2, first GPU pass on ONE GPU: parity tests, q3 at a small scale factor, then the SF100 headline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2a_pytest_gpu.txt
echo "--- pytest done"; tail -5 gpurun_out/r2a_pytest_gpu.txt
timeout 600 python bench.py --sf 1 --ref-sf 1 --steps 3 --extra-q6 0 > gpurun_out/r2a_bench_sf1.json 2> gpurun_out/r2a_bench_sf1.err
echo "--- sf1 rc=$?"; tail -c 1500 gpurun_out/r2a_bench_sf1.json; tail -5 gpurun_out/r2a_bench_sf1.err
timeout 900 python bench.py --steps 5 --extra-q6 0 > gpurun_out/r2a_bench_sf100.json 2> gpurun_out/r2a_bench_sf100.err
echo "--- sf100 rc=$?"; tail -c 3000 gpurun_out/r2a_bench_sf100.json; tail -8 gpurun_out/r2a_bench_sf100.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
free -g | head -2
nproc
