#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py -x -q -s -m gpu > gpurun_out/r2t_tests.log 2>&1; echo tests rc=$?; grep -E "picked|passed|failed" gpurun_out/r2t_tests.log | tail -4
for c in 4 2 1; do
python bench.py --config $c --no-e2e --no-cpu-baseline > gpurun_out/r2t_bench_c$c.json 2> gpurun_out/r2t_bench_c$c.err; echo bench$c rc=$?
done
python - <<'PY'
import json
for c in (4,2,1):
    d=json.load(open(f"gpurun_out/r2t_bench_c{c}.json"))
    print(c, round(d["ms_per_step"],3), d["kernel_ms_rank0"], d["scan_long_build"], round(d["roofline"]["frac"],4), round(d["roofline"]["whole_path_frac"],4))
PY
