#!/bin/bash
# pass P: full GPU tests + q3 SF100 bench (filter fused into the probe) + the same with the fusion off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.txt 2>&1; echo "--- pytest rc=$?"; tail -4 gpurun_out/r2p_pytest.txt
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --extra-q6 0 --cpu-baseline 0 --check ${CHECK:-0} > gpurun_out/r2p_$name.json 2> gpurun_out/r2p_$name.err; echo "--- $name rc=$?"; tail -2 gpurun_out/r2p_$name.err
  python - "$name" <<'PY'
import json, sys
f = "r2p_" + sys.argv[1]
try:
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, round(d["value"] / 1e9, 3), "G rows/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["ms_per_step"], 1), d["config"].get("checked"))
    for o in d["operators"]: print("  op", o["name"], round(o["ms_per_step"], 3), o.get("rows_in"), o["rows_out"])
    for k in d["kernels"][:8]: print("  k", k["name"], round(k["ms_per_step"], 3), round(k["launches_per_step"], 1))
except Exception as e:
    print(f, "ERR", e)
PY
}
CHECK=1 run base A=1
run nopred B2_JOIN_NO_PRED_FUSION=1
