#!/bin/bash
# round 2, fourth GPU call: per-warp work distribution (no block barrier) A/B; run_loop + exchange tests
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2d_suite.log 2>&1; echo suite rc=$?; tail -6 gpurun_out/r2d_suite.log
run() { # tag env lib
  SDM_SCAN=$2 SDM_LIB=$3 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2d_$1.json 2> gpurun_out/r2d_$1.err; echo $1 rc=$?
}
L=$PWD/eao-slam_b200/lib
run head3 lane3 $L/ab/libsdm_head.so
run warp3 lane3 $L/libsdm_b200.so
run head2 lane2 $L/ab/libsdm_head.so
run warp2 lane2 $L/libsdm_b200.so
run head3b lane3 $L/ab/libsdm_head.so
run warp3b lane3 $L/libsdm_b200.so
python - <<'PY'
import json
for n in ("head3","warp3","head2","warp2","head3b","warp3b"):
    try:
        d=json.load(open(f"gpurun_out/r2d_{n}.json"))
        print(n, round(d["ms_per_step"],3), d["kernel_ms_rank0"], d.get("fused_per_step_rank0"), d.get("checked_per_step_rank0"), d.get("scan_generation"))
    except Exception as e:
        print(n, "failed", e)
PY
