#!/bin/bash
# round 2 evidence on ONE GPU: launch list of one q3 step + full captures of the dominant kernels (SF100 q3 bench).
# The .ncu-rep files of the VM kernels are large (source import): they are reduced to CSV / text pages ON THE BOX and removed.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --cpu-baseline 0 --check 0 --extra-q6 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 390 -c 130 --csv --log-file gpurun_out/r2_launches_q3.csv $BENCH > gpurun_out/r2_launches_q3.log 2>&1
echo "--- launches rc=$?"
for K in join_probe_distinct1_kernel simple_filter_ids_kernel radix_agg_kernel part_scatter2_kernel gather_fixed_kernel radix_rows_kernel join_build_kernel filter_staged_kernel; do
  SRC="--import-source on"; case $K in filter_staged_kernel|radix_rows_kernel) SRC="";; esac
  SKIP=40; case $K in radix_agg_kernel|radix_rows_kernel) SKIP=3;; part_scatter2_kernel) SKIP=6;; gather_fixed_kernel) SKIP=100;; join_build_kernel) SKIP=7;; filter_staged_kernel) SKIP=4;; esac
  timeout 900 ncu --set full --clock-control none $SRC -k regex:$K -s $SKIP -c 1 -o gpurun_out/r2_ncu_$K -f $BENCH > gpurun_out/r2_ncu_$K.log 2>&1
  echo "--- $K rc=$?"
  if [ -f gpurun_out/r2_ncu_$K.ncu-rep ]; then
    ncu -i gpurun_out/r2_ncu_$K.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${K}_raw.csv 2>/dev/null
    ncu -i gpurun_out/r2_ncu_$K.ncu-rep --page details > gpurun_out/r2_ncu_${K}_details.txt 2>/dev/null
    if [ -n "$SRC" ]; then ncu -i gpurun_out/r2_ncu_$K.ncu-rep --page source --csv > gpurun_out/r2_ncu_${K}_source.csv 2>/dev/null; fi
    SZ=$(stat -c %s gpurun_out/r2_ncu_$K.ncu-rep); if [ "$SZ" -gt 8000000 ]; then rm gpurun_out/r2_ncu_$K.ncu-rep; fi
  fi
done
du -sh gpurun_out; ls -la gpurun_out | head -40
