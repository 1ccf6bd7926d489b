#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_memory_gpu.py tests/test_shuffle_format_gpu.py tests/test_parquet_gpu.py::test_parquet_chunked_reader_row_group_chunks tests/test_parquet_gpu.py::test_parquet_corrupt_page_header_is_rejected tests/test_execs_gpu.py tests/test_relational_gpu.py -m gpu -q --tb=long 2>&1 | tail -150 > gpurun_out/r2e_pytest_subset.txt
echo "--- subset done"; tail -15 gpurun_out/r2e_pytest_subset.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2e_pytest_gpu.txt
echo "--- pytest done"; tail -8 gpurun_out/r2e_pytest_gpu.txt
run() { name=$1; shift; env "$@" timeout 900 python bench.py --steps 5 --extra-q6 0 --cpu-baseline 0 > gpurun_out/r2e_$name.json 2> gpurun_out/r2e_$name.err; echo "--- $name rc=$?"; tail -3 gpurun_out/r2e_$name.err; }
run sf100 B2_X=1
run sf100_nofusion B2_NO_FILTER_FUSION=1
python - <<'PY'
import json
for f in ["r2e_sf100","r2e_sf100_nofusion"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, round(d["value"]/1e9,3), "G rows/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],1), d["config"].get("check_s"), d["config"].get("checked"))
    for o in d["operators"]: print("  op", o["name"], round(o["ms_per_step"],3), o.get("rows_in"), o["rows_out"], "frac", round(o.get("hbm_frac",0),4))
    for k in d["kernels"]: print("  k", k["name"], round(k["ms_per_step"],3), round(k["launches_per_step"],1), round(k["share"],3), round(k.get("alg_GBps",0),1))
PY
