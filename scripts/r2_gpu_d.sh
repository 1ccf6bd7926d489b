#!/bin/bash
# round 2, 2-GPU pass: parity tests (incl. the multi-rank exchange check through pytest), the multi-rank script, the strong-scaled SF100 q3 at N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2d_pytest_gpu.txt
echo "--- pytest done"; tail -12 gpurun_out/r2d_pytest_gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/exchange_check.py > gpurun_out/r2d_exchange_check.txt 2>&1
echo "--- exchange_check rc=$?"; tail -15 gpurun_out/r2d_exchange_check.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --cpu-baseline 0 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
echo "--- bench n2 rc=$?"; tail -5 gpurun_out/r2d_bench_n2.err
python - <<'PY'
import json
for f in ["r2d_bench_n2"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, round(d["value"]/1e9,3), "G rows/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],1), d["config"].get("check_s"), d["config"].get("checked"))
    print(" exchange", d.get("exchange"))
    for o in d["operators"]: print("  op", o["name"], round(o["ms_per_step"],3), o.get("rows_in"), o["rows_out"], "frac", round(o.get("hbm_frac",0),4))
    for k in d["kernels"]: print("  k", k["name"], round(k["ms_per_step"],3), round(k["launches_per_step"],1), round(k["share"],3), round(k.get("alg_GBps",0),1))
PY
