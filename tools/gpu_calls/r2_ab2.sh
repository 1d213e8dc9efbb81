#!/bin/bash
# round 2, second GPU call: whole GPU suite under the third-generation loop, register / unroll variants, ncu capture
mkdir -p gpurun_out
SDM_SCAN=lane3 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_suite_lane3.log 2>&1; echo suite-lane3 rc=$?; tail -4 gpurun_out/r2b_suite_lane3.log
run() { # tag env lib
  SDM_SCAN=$2 SDM_LIB=$3 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2b_$1.json 2> gpurun_out/r2b_$1.err; echo $1 rc=$?
}
L=$PWD/eao-slam_b200/lib
run lane2 lane2 $L/libsdm_b200.so
run l3m10u2 lane3 $L/libsdm_b200.so
run l3m10u1 lane3 $L/ab/libsdm_l3m10u1.so
run l3m9u2 lane3 $L/ab/libsdm_l3m9u2.so
run l3m8u2 lane3 $L/ab/libsdm_l3m8u2.so
run lane2b lane2 $L/libsdm_b200.so
python - <<'PY'
import json
for n in ("lane2","l3m10u2","l3m10u1","l3m9u2","l3m8u2","lane2b"):
    try:
        d=json.load(open(f"gpurun_out/r2b_{n}.json"))
        print(n, round(d["ms_per_step"],3), d["kernel_ms_rank0"], d.get("fused_per_step_rank0"), d.get("checked_per_step_rank0"), d.get("scan_generation"))
    except Exception as e:
        print(n, "failed", e)
PY
SDM_SCAN=lane3 ncu --set full --clock-control none --import-source on -k regex:"k_pass1_lane" -s 3 -c 1 -f -o gpurun_out/prof_r2b_lane3 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-hot-spin > gpurun_out/ncu_r2b_lane3.log 2>&1; echo ncu-lane3 rc=$?
SDM_SCAN=lane2 ncu --set full --clock-control none --import-source on -k regex:"k_pass1_lane" -s 3 -c 1 -f -o gpurun_out/prof_r2b_lane2 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-hot-spin > gpurun_out/ncu_r2b_lane2.log 2>&1; echo ncu-lane2 rc=$?
ls -la gpurun_out/*.ncu-rep
