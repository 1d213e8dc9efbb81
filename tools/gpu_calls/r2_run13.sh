#!/bin/bash
# single GPU: line-fit tests incl. the real-ED-mask loop, line-fit timing on VGA chains, default bench with e2e_online
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linefit.py -x -q -s > gpurun_out/r2n_linefit.log 2>&1; echo linefit rc=$?; tail -8 gpurun_out/r2n_linefit.log
timeout 600 python tools/linefit_bench.py > gpurun_out/r2n_linefit_bench.json 2> gpurun_out/r2n_linefit_bench.err; echo lfbench rc=$?; cat gpurun_out/r2n_linefit_bench.json; tail -3 gpurun_out/r2n_linefit_bench.err
python bench.py > gpurun_out/r2n_bench_c2.json 2> gpurun_out/r2n_bench_c2.err; echo bench rc=$?; tail -3 gpurun_out/r2n_bench_c2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2n_bench_c2.json"))
print(round(d["ms_per_step"],3), d["kernel_ms_rank0"], "value %.3e"%d["value"])
for k in ("e2e","e2e_class","e2e_online"):
    if k in d: print("   ", k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in d[k].items() if a not in ("api","note","sample")})
PY
