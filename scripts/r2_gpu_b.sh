#!/bin/bash
# round 2, GPU pass: parity tests + the SF100 headline (no CPU arm) + A/B of the TMA-staged filter and the Bloom prefilter
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2b_pytest_gpu.txt
echo "--- pytest done"; tail -5 gpurun_out/r2b_pytest_gpu.txt
timeout 900 python bench.py --steps 5 --extra-q6 0 --cpu-baseline 0 > gpurun_out/r2b_bench_sf100.json 2> gpurun_out/r2b_bench_sf100.err
echo "--- sf100 rc=$?"; tail -3 gpurun_out/r2b_bench_sf100.err
B2_FILTER_NO_TMA=1 timeout 900 python bench.py --steps 5 --extra-q6 0 --cpu-baseline 0 --check 0 > gpurun_out/r2b_bench_sf100_notma.json 2> gpurun_out/r2b_bench_sf100_notma.err
B2_JOIN_NO_BLOOM=1 timeout 900 python bench.py --steps 5 --extra-q6 0 --cpu-baseline 0 --check 0 > gpurun_out/r2b_bench_sf100_nobloom.json 2> gpurun_out/r2b_bench_sf100_nobloom.err
python - <<'PY'
import json
for f in ["r2b_bench_sf100","r2b_bench_sf100_notma","r2b_bench_sf100_nobloom"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, round(d["value"]/1e9,3), "G rows/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],1), d["config"].get("check_s"), d["config"].get("checked"))
    for o in d["operators"]: print("  op", o["name"], round(o["ms_per_step"],3), o.get("rows_in"), o["rows_out"], "frac", round(o.get("hbm_frac",0),4))
    for k in d["kernels"]: print("  k", k["name"], round(k["ms_per_step"],3), round(k["launches_per_step"],1), round(k["share"],3), round(k.get("alg_GBps",0),1))
PY
