#!/bin/bash
# pass S: relational / aggregation tests + q3 SF100 bench with cache-policy hints in the scatter kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_relational_gpu.py tests/test_execs_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > gpurun_out/r2s_pytest.txt 2>&1; echo "--- pytest rc=$?"; tail -3 gpurun_out/r2s_pytest.txt
timeout 600 python bench.py --steps 5 --extra-q6 0 --cpu-baseline 0 --check 1 > gpurun_out/r2s_base.json 2> gpurun_out/r2s_base.err; echo "--- base rc=$?"; tail -2 gpurun_out/r2s_base.err
python - <<'PY'
import json
f = "r2s_base"
try:
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, round(d["value"] / 1e9, 3), "G rows/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["ms_per_step"], 1), d["config"].get("checked"))
    for o in d["operators"]: print("  op", o["name"], round(o["ms_per_step"], 3))
    for k in d["kernels"][:12]: print("  k", k["name"], round(k["ms_per_step"], 3), round(k["launches_per_step"], 1))
except Exception as e:
    print(f, "ERR", e)
PY
