#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2g_pytest_gpu.txt
echo "--- pytest done"; tail -8 gpurun_out/r2g_pytest_gpu.txt
timeout 300 python scripts/xchg_local_probe.py > gpurun_out/r2g_xchg_local.txt 2>&1; tail -12 gpurun_out/r2g_xchg_local.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xchg_scatter_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_xchg_scatter_kernel -f python scripts/xchg_local_probe.py > gpurun_out/r2_ncu_xchg.log 2>&1
ncu -i gpurun_out/r2_ncu_xchg_scatter_kernel.ncu-rep --page raw --csv > gpurun_out/r2_ncu_xchg_scatter_kernel_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_ncu_xchg_scatter_kernel.ncu-rep --page details > gpurun_out/r2_ncu_xchg_scatter_kernel_details.txt 2>/dev/null
ncu -i gpurun_out/r2_ncu_xchg_scatter_kernel.ncu-rep --page source --csv > gpurun_out/r2_ncu_xchg_scatter_kernel_source.csv 2>/dev/null
bash scripts/r2_ncu.sh
