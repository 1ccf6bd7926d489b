#!/bin/bash
# source-level capture of the small-page snappy kernel on a 6M-row q6 file (one launch)
ncu --set full --import-source on --clock-control none -k regex:snappy_kernel -s 7 -c 1 -o gpurun_out/snappy_pipe -f python bench.py --rows 12000000 --steps 1 --warmup 3 --cpu-baseline 0 > gpurun_out/ncu_snappy.log 2>&1
ncu -i gpurun_out/snappy_pipe.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/snappy_pipe_source.csv 2>/dev/null
ncu -i gpurun_out/snappy_pipe.ncu-rep --page raw --csv > gpurun_out/snappy_pipe_raw.csv 2>/dev/null
