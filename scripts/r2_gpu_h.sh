#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --extra-q6 0 --cpu-baseline 0 > gpurun_out/r2h_sf100.json 2> gpurun_out/r2h_sf100.err; echo "--- sf100 rc=$?"; tail -3 gpurun_out/r2h_sf100.err
python - <<'PY'
import json
for f in ["r2h_sf100"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, round(d["value"]/1e9,3), "G rows/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],1), d["config"].get("check_s"), d["config"].get("checked"))
    for o in d["operators"]: print("  op", o["name"], round(o["ms_per_step"],3), o.get("rows_in"), o["rows_out"], "frac", round(o.get("hbm_frac",0),4))
    for k in d["kernels"]: print("  k", k["name"], round(k["ms_per_step"],3), round(k["launches_per_step"],1), round(k["share"],3), round(k.get("alg_GBps",0),1))
PY
timeout 300 python scripts/xchg_local_probe.py > gpurun_out/r2h_xchg_local.txt 2>&1; tail -6 gpurun_out/r2h_xchg_local.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xchg_scatter_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_xchg_scatter_kernel -f python scripts/xchg_local_probe.py > gpurun_out/r2_ncu_xchg.log 2>&1
ncu -i gpurun_out/r2_ncu_xchg_scatter_kernel.ncu-rep --page raw --csv > gpurun_out/r2_ncu_xchg_scatter_kernel_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_ncu_xchg_scatter_kernel.ncu-rep --page details > gpurun_out/r2_ncu_xchg_scatter_kernel_details.txt 2>/dev/null
rm -f gpurun_out/r2_ncu_xchg_scatter_kernel.ncu-rep
bash scripts/r2_ncu.sh
