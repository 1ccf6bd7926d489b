#!/bin/bash
# full-set captures of the two heaviest kernels of the SF10 q6 step: snappy_kernel (main launch) and aggregate_kernel.
# Only text summaries come back (the .ncu-rep of the VM kernel with all its handler instantiations is > 60 MB).
cd "$(dirname "$0")/.."
T=/tmp/ncu_r1; mkdir -p $T gpurun_out
ncu --set full --import-source on --clock-control none -k regex:^snappy_kernel -s 6 -c 1 -o $T/snappy -f python bench.py --steps 1 --warmup 3 --cpu-baseline 0 > gpurun_out/ncu_snappy.log 2>&1
ncu -i $T/snappy.ncu-rep --page source --csv --print-source cuda,sass > $T/snappy_source.csv 2>/dev/null
ncu -i $T/snappy.ncu-rep --page raw --csv > gpurun_out/r1_snappy_sf10_raw.csv 2>/dev/null
python scripts/ncu_lines.py $T/snappy_source.csv 60 > gpurun_out/r1_snappy_sf10_lines.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:aggregate_kernel -s 3 -c 1 -o $T/agg -f python bench.py --steps 1 --warmup 3 --cpu-baseline 0 > gpurun_out/ncu_agg.log 2>&1
ncu -i $T/agg.ncu-rep --page source --csv --print-source cuda,sass > $T/agg_source.csv 2>/dev/null
ncu -i $T/agg.ncu-rep --page raw --csv > gpurun_out/r1_agg_sf10_raw.csv 2>/dev/null
python scripts/ncu_lines.py $T/agg_source.csv 60 > gpurun_out/r1_agg_sf10_lines.txt 2>&1
ls -la $T gpurun_out
