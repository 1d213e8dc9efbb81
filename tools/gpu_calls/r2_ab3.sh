#!/bin/bash
# round 2, third GPU call: 32-byte scan texel (one sector per visited column) - parity + A/B at 10 / 12 blocks per SM
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2c_suite.log 2>&1; echo suite rc=$?; tail -4 gpurun_out/r2c_suite.log
run() { # tag env lib
  SDM_SCAN=$2 SDM_LIB=$3 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c_$1.json 2> gpurun_out/r2c_$1.err; echo $1 rc=$?
}
L=$PWD/eao-slam_b200/lib
run lane2 lane2 $L/libsdm_b200.so
run l3m10 lane3 $L/libsdm_b200.so
run l3m12 lane3 $L/ab/libsdm_l3m12.so
run l3m10b lane3 $L/libsdm_b200.so
python - <<'PY'
import json
for n in ("lane2","l3m10","l3m12","l3m10b"):
    try:
        d=json.load(open(f"gpurun_out/r2c_{n}.json"))
        print(n, round(d["ms_per_step"],3), d["kernel_ms_rank0"], d.get("fused_per_step_rank0"), d.get("checked_per_step_rank0"), d.get("scan_generation"))
    except Exception as e:
        print(n, "failed", e)
PY
SDM_SCAN=lane3 ncu --set full --clock-control none --import-source on -k regex:"k_pass1_lane" -s 3 -c 1 -f -o gpurun_out/prof_r2c_lane3 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-hot-spin > gpurun_out/ncu_r2c_lane3.log 2>&1; echo ncu-lane3 rc=$?
