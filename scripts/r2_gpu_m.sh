#!/bin/bash
# pass M (1 GPU): tests + q3 SF100 bench + local exchange probe + ncu of radix_agg / xchg_scatter
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.txt 2>&1; echo "--- pytest rc=$?"; tail -4 gpurun_out/r2m_pytest.txt
timeout 600 python bench.py --steps 5 --extra-q6 0 --cpu-baseline 0 --check 1 > gpurun_out/r2m_base.json 2> gpurun_out/r2m_base.err; echo "--- base rc=$?"; tail -2 gpurun_out/r2m_base.err
python - <<'PY'
import json
f = "r2m_base"
try:
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, round(d["value"] / 1e9, 3), "G rows/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["ms_per_step"], 1), d["config"].get("checked"))
    for o in d["operators"]: print("  op", o["name"], round(o["ms_per_step"], 3))
    for k in d["kernels"][:14]: print("  k", k["name"], round(k["ms_per_step"], 3), round(k["launches_per_step"], 1))
except Exception as e:
    print(f, "ERR", e)
PY
BENCH="python bench.py --steps 1 --warmup 3 --cpu-baseline 0 --check 0 --extra-q6 0"
for K in radix_agg_kernel; do
  timeout 600 ncu --set full --clock-control none -k regex:$K -s 3 -c 1 -o gpurun_out/r2m_ncu_$K -f $BENCH > gpurun_out/r2m_ncu_$K.log 2>&1
  echo "--- $K rc=$?"
  if [ -f gpurun_out/r2m_ncu_$K.ncu-rep ]; then
    ncu -i gpurun_out/r2m_ncu_$K.ncu-rep --page raw --csv > gpurun_out/r2m_ncu_${K}_raw.csv 2>/dev/null
    rm gpurun_out/r2m_ncu_$K.ncu-rep
  fi
done
python scripts/ncu_summarize.py gpurun_out/r2m_ncu_*_raw.csv
