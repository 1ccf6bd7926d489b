#!/bin/bash
# pass N: sort tests (small-sort path) + Bloom filter size sweep with the new probe kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_relational_gpu.py tests/test_execs_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > gpurun_out/r2n_pytest.txt 2>&1; echo "--- pytest rc=$?"; tail -4 gpurun_out/r2n_pytest.txt
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --extra-q6 0 --cpu-baseline 0 --check ${CHECK:-0} > gpurun_out/r2n_$name.json 2> gpurun_out/r2n_$name.err; echo "--- $name rc=$?"; tail -2 gpurun_out/r2n_$name.err
  python - "$name" <<'PY'
import json, sys
f = "r2n_" + sys.argv[1]
try:
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, round(d["value"] / 1e9, 3), "G rows/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["ms_per_step"], 1), d["config"].get("checked"))
    for o in d["operators"][:5]: print("  op", o["name"], round(o["ms_per_step"], 3))
    for k in d["kernels"][:6]: print("  k", k["name"], round(k["ms_per_step"], 3), round(k["launches_per_step"], 1))
except Exception as e:
    print(f, "ERR", e)
PY
}
CHECK=1 run base A=1
run bloom8 B2_JOIN_BLOOM_BITS=8
run bloom4 B2_JOIN_BLOOM_BITS=4
run bloom8p B2_JOIN_BLOOM_BITS=8 B2_JOIN_BLOOM_PERSIST=1
run bloom16p B2_JOIN_BLOOM_BITS=16 B2_JOIN_BLOOM_PERSIST=1
