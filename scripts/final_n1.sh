#!/bin/bash
# round-end evidence on one GPU: parity tests, both bench arms, launch list, operator micro-benchmarks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/final_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --cpu-baseline 0 > gpurun_out/final_ncu_bench.log 2>&1
timeout 300 python scripts/operators_bench.py > gpurun_out/final_operators.txt 2>&1
timeout 300 python scripts/q1_columns.py > gpurun_out/final_q1.txt 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.txt 2>&1
cat gpurun_out/final_pytest_gpu.txt gpurun_out/final_smoke.txt; tail -c 600 gpurun_out/final_bench_n1.json; tail -c 400 gpurun_out/final_bench_reference.json; tail -5 gpurun_out/final_q1.txt
