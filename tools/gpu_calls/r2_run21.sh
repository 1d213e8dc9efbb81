#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_run_loop.py -x -q -m gpu > gpurun_out/r2y_runloop.log 2>&1; echo runloop rc=$?; tail -3 gpurun_out/r2y_runloop.log
python bench.py --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo bench rc=$?; tail -3 gpurun_out/r2y_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2y_bench.json"))
print(round(d["ms_per_step"],3), d["kernel_ms_rank0"])
for k in ("e2e","e2e_blocks","e2e_points_blocks","e2e_class"):
    if k in d: print("   ", k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in d[k].items() if a not in ("api","note","sample")})
PY
