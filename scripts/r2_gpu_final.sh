#!/bin/bash
# round 2 evidence at HEAD on ONE GPU: GPU tests, smoke(), the default bench line (as the driver runs it), the reference arm,
# the per-operator bench, then the ncu launch list + full captures (scripts/r2_ncu.sh)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest_gpu.txt 2>&1; echo "--- pytest rc=$?"; tail -3 gpurun_out/r2f_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2f_smoke.txt 2>&1; echo "--- smoke rc=$?"; tail -2 gpurun_out/r2f_smoke.txt
timeout 900 python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; echo "--- bench default rc=$?"; tail -2 gpurun_out/r2f_bench_default.err
timeout 900 python bench.py --impl reference > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo "--- bench reference rc=$?"; tail -2 gpurun_out/r2f_bench_reference.err
python - <<'PY'
import json
for f in ["r2f_bench_default", "r2f_bench_reference"]:
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["value"] / 1e9, 3), "G rows/s", round(d.get("ms_per_step", 0), 2), "ms; e2e", d.get("e2e", {}).get("ms_per_step"), d["config"].get("checked"), d.get("clocks"), d.get("cpu_baseline"))
        print("  roofline", d.get("roofline"))
        for o in d.get("operators", []): print("  op", o["name"], round(o["ms_per_step"], 3), o.get("rows_in"), o["rows_out"], round(o.get("hbm_frac", 0), 4))
        for k in d.get("kernels", []): print("  k", k["name"], round(k["ms_per_step"], 3), round(k["launches_per_step"], 1), round(k.get("alg_GBps", 0), 1))
        if "extra" in d: print("  extra", json.dumps(d["extra"])[:600])
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 python scripts/operators_bench.py > gpurun_out/r2f_operators_bench.txt 2>&1; echo "--- operators rc=$?"; tail -22 gpurun_out/r2f_operators_bench.txt
cp gpurun_out/operators_bench.json gpurun_out/r2f_operators_bench.json 2>/dev/null
bash scripts/r2_ncu.sh
python scripts/ncu_summarize.py gpurun_out/r2_ncu_*_raw.csv --json gpurun_out/r2_ncu_q3_traffic.json > gpurun_out/r2_ncu_summary.txt 2>&1; tail -40 gpurun_out/r2_ncu_summary.txt
