#!/bin/bash
# round 2, N-GPU pass (N = number of visible GPUs): multi-rank exchange parity + the strong-scaled SF100 q3 bench at N
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/exchange_check.py > gpurun_out/r2l_exchange_check_n$N.txt 2>&1
echo "--- exchange_check rc=$?"; tail -12 gpurun_out/r2l_exchange_check_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 5 --cpu-baseline 0 --extra-q6 0 > gpurun_out/r2l_bench_n$N.json 2> gpurun_out/r2l_bench_n$N.err
echo "--- bench n$N rc=$?"; tail -5 gpurun_out/r2l_bench_n$N.err
python - $N <<'PY'
import json, sys
f = "r2l_bench_n" + sys.argv[1]
try:
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, round(d["value"] / 1e9, 3), "G rows/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["ms_per_step"], 1), d["config"].get("check_s"), d["config"].get("checked"))
    print(" exchange", d.get("exchange"))
    for o in d["operators"]: print("  op", o["name"], round(o["ms_per_step"], 3), o.get("rows_in"), o["rows_out"])
    for k in d["kernels"][:14]: print("  k", k["name"], round(k["ms_per_step"], 3), round(k["launches_per_step"], 1))
except Exception as e:
    print(f, "ERR", e)
PY
