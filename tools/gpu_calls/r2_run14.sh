#!/bin/bash
# single GPU: warp-per-chain line fit (tests + timing), context creation time
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linefit.py -x -q -s > gpurun_out/r2o_linefit.log 2>&1; echo linefit rc=$?; tail -8 gpurun_out/r2o_linefit.log
timeout 600 python tools/linefit_bench.py > gpurun_out/r2o_linefit_bench.json 2> gpurun_out/r2o_linefit_bench.err; echo lfbench rc=$?; cat gpurun_out/r2o_linefit_bench.json; tail -3 gpurun_out/r2o_linefit_bench.err
python - <<'PY' 2>&1 | tee gpurun_out/r2o_create.log
import sys, time
sys.path.insert(0, "eao-slam_b200/python")
from sdmb200 import api
api.load()
for n in (32, 65, 65, 200):
    t = time.perf_counter()
    ctx = api.Context(width=640, height=480, max_keyframes=n)
    t1 = time.perf_counter()
    ctx.close()
    print(f"sdm_create {n} slots VGA: {1e3*(t1-t):.1f} ms, destroy {1e3*(time.perf_counter()-t1):.1f} ms")
PY
