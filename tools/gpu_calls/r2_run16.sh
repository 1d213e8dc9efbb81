#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_suite.log 2>&1; echo suite rc=$?; tail -3 gpurun_out/r2q_suite.log
python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo bench rc=$?
python bench.py --config 4 --no-e2e --no-cpu-baseline > gpurun_out/r2q_bench_c4.json 2> gpurun_out/r2q_bench_c4.err; echo bench4 rc=$?
python - <<'PY'
import json
for n in ("r2q_bench","r2q_bench_c4"):
    d=json.load(open(f"gpurun_out/{n}.json"))
    print(n, d["ms_per_step"], d["kernel_ms_rank0"], d["roofline"]["whole_path_frac_with_compaction"])
PY
