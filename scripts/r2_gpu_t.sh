#!/bin/bash
# pass T: evidence refresh at HEAD after the cache-policy hints in part_scatter2: default bench line, reference arm, launch list,
# full capture of part_scatter2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; echo "--- bench default rc=$?"; tail -2 gpurun_out/r2f_bench_default.err
timeout 900 python bench.py --impl reference > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo "--- bench reference rc=$?"
python - <<'PY'
import json
for f in ["r2f_bench_default", "r2f_bench_reference"]:
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["value"] / 1e9, 3), "G rows/s", round(d.get("ms_per_step", 0), 2), "ms; e2e", d.get("e2e", {}).get("ms_per_step"), d["config"].get("checked"), d.get("clocks"))
        for o in d.get("operators", []): print("  op", o["name"], round(o["ms_per_step"], 3))
        for k in d.get("kernels", []): print("  k", k["name"], round(k["ms_per_step"], 3), round(k["launches_per_step"], 1), round(k.get("alg_GBps", 0), 1))
    except Exception as e:
        print(f, "ERR", e)
PY
BENCH="python bench.py --steps 1 --warmup 3 --cpu-baseline 0 --check 0 --extra-q6 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 390 -c 130 --csv --log-file gpurun_out/r2_launches_q3.csv $BENCH > gpurun_out/r2_launches_q3.log 2>&1; echo "--- launches rc=$?"
K=part_scatter2_kernel
timeout 600 ncu --set full --clock-control none -k regex:$K -s 6 -c 1 -o gpurun_out/r2_ncu_$K -f $BENCH > gpurun_out/r2_ncu_$K.log 2>&1; echo "--- $K rc=$?"
ncu -i gpurun_out/r2_ncu_$K.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${K}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_ncu_$K.ncu-rep --page details > gpurun_out/r2_ncu_${K}_details.txt 2>/dev/null
rm -f gpurun_out/r2_ncu_$K.ncu-rep
python scripts/ncu_summarize.py gpurun_out/r2_ncu_${K}_raw.csv
