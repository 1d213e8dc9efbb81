#!/bin/bash
# round 2, first GPU call: third-generation scan loop - parity under the whole GPU suite, then A/B timing
mkdir -p gpurun_out
python -m pytest tests/test_gpu_edge_cases.py -x -q -m gpu -k "generations" > gpurun_out/r2a_gen.log 2>&1; echo gen-test rc=$?; tail -3 gpurun_out/r2a_gen.log
SDM_SCAN=lane3 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_edge_cases.py::test_scan_generations_agree > gpurun_out/r2a_suite_lane3.log 2>&1; echo suite-lane3 rc=$?; tail -5 gpurun_out/r2a_suite_lane3.log
for v in lane2 lane3; do
  SDM_SCAN=$v python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_$v.json 2> gpurun_out/r2a_$v.err; echo $v rc=$?
done
SDM_SCAN=lane3 SDM_LIB=$PWD/eao-slam_b200/lib/ab/libsdm_l3m10.so python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_lane3_m10.json 2> gpurun_out/r2a_lane3_m10.err; echo lane3-m10 rc=$?
python - <<'PY'
import json
for n in ("lane2","lane3","lane3_m10"):
    try:
        d=json.load(open(f"gpurun_out/r2a_{n}.json"))
        print(n, round(d["ms_per_step"],3), d["kernel_ms_rank0"], d.get("fused_per_step_rank0"), d.get("checked_per_step_rank0"), d.get("scan_generation"))
    except Exception as e:
        print(n, "failed", e)
PY
