#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2c_pytest_gpu.txt
echo "--- pytest done"; tail -8 gpurun_out/r2c_pytest_gpu.txt
run() { name=$1; shift; env "$@" timeout 900 python bench.py --steps 5 --extra-q6 0 --cpu-baseline 0 > gpurun_out/r2c_$name.json 2> gpurun_out/r2c_$name.err; echo "--- $name rc=$?"; tail -3 gpurun_out/r2c_$name.err; }
run sf100 B2_X=1
run sf100_notma B2_FILTER_NO_TMA=1
run sf100_noradix B2_AGG_NO_RADIX=1
python - <<'PY'
import json
for f in ["r2c_sf100","r2c_sf100_notma","r2c_sf100_noradix"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, round(d["value"]/1e9,3), "G rows/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],1), d["config"].get("check_s"), d["config"].get("checked"))
    for o in d["operators"]: print("  op", o["name"], round(o["ms_per_step"],3), o.get("rows_in"), o["rows_out"], "frac", round(o.get("hbm_frac",0),4))
    for k in d["kernels"]: print("  k", k["name"], round(k["ms_per_step"],3), round(k["launches_per_step"],1), round(k["share"],3), round(k.get("alg_GBps",0),1))
PY
