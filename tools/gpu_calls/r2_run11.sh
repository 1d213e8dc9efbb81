#!/bin/bash
# single GPU: line-fit tests first, then the whole suite, smoke, default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linefit.py -x -q -s > gpurun_out/r2l_linefit.log 2>&1; echo linefit rc=$?; tail -15 gpurun_out/r2l_linefit.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2l_suite.log 2>&1; echo suite rc=$?; tail -4 gpurun_out/r2l_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.log 2>&1; echo smoke rc=$?
python bench.py > gpurun_out/r2l_bench_c2.json 2> gpurun_out/r2l_bench_c2.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2l_bench_c2.json"))
print(round(d["ms_per_step"],3), d["kernel_ms_rank0"], "value %.3e"%d["value"], d["roofline"])
for k in ("e2e","e2e_class","e2e_point_export","e2e_image_in_points_out","cpu_baseline"):
    if k in d: print("   ", k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in d[k].items() if a not in ("api","note","sample")})
PY
