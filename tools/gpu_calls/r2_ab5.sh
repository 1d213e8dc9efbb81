#!/bin/bash
# round 2, fifth GPU call: suite with the reworked class shim; A/B: intensity prefetch, candidate tile shapes
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2e_suite.log 2>&1; echo suite rc=$?; tail -8 gpurun_out/r2e_suite.log
run() { # tag env lib
  SDM_SCAN=$2 SDM_LIB=$3 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2e_$1.json 2> gpurun_out/r2e_$1.err; echo $1 rc=$?
}
L=$PWD/eao-slam_b200/lib
run base lane3 $L/libsdm_b200.so
run pfim lane3 $L/ab/libsdm_pfim.so
run t16 lane3 $L/ab/libsdm_t16.so
run t8 lane3 $L/ab/libsdm_t8.so
run base2 lane3 $L/libsdm_b200.so
python - <<'PY'
import json
for n in ("base","pfim","t16","t8","base2"):
    try:
        d=json.load(open(f"gpurun_out/r2e_{n}.json"))
        print(n, round(d["ms_per_step"],3), d["kernel_ms_rank0"], d.get("fused_per_step_rank0"), d.get("checked_per_step_rank0"), d.get("scan_generation"))
    except Exception as e:
        print(n, "failed", e)
PY
