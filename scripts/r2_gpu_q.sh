#!/bin/bash
# pass Q: Bloom geometry sweep (bits per key x bits set per key) on the q3 SF100 step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --extra-q6 0 --cpu-baseline 0 --check ${CHECK:-0} > gpurun_out/r2q_$name.json 2> gpurun_out/r2q_$name.err; echo "--- $name rc=$?"; tail -2 gpurun_out/r2q_$name.err
  python - "$name" <<'PY'
import json, sys
f = "r2q_" + sys.argv[1]
try:
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, round(d["value"] / 1e9, 3), "G rows/s", round(d["ms_per_step"], 2), "ms;", d["config"].get("checked"))
    for k in d["kernels"][:6]:
        if "join" in k["name"]: print("  k", k["name"], round(k["ms_per_step"], 3), round(k["launches_per_step"], 1))
except Exception as e:
    print(f, "ERR", e)
PY
}
CHECK=1 run b16k2 A=1
CHECK=1 run b8k3 B2_JOIN_BLOOM_BITS=8 B2_JOIN_BLOOM_K=3
run b8k4 B2_JOIN_BLOOM_BITS=8 B2_JOIN_BLOOM_K=4
run b16k3 B2_JOIN_BLOOM_BITS=16 B2_JOIN_BLOOM_K=3
run b32k2 B2_JOIN_BLOOM_BITS=32
