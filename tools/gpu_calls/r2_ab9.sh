#!/bin/bash
# chunk hand-out of the scan kernel: per-warp pieces (previous default) vs block chunks through a shared ring (no barrier), configs 2 and 4;
# then the edge-case + parity tests on the new default
mkdir -p gpurun_out
for c in 2 4; do
for lib in eao-slam_b200/lib/libsdm_b200.so eao-slam_b200/lib/variants/libsdm_warp_pieces.so eao-slam_b200/lib/libsdm_b200.so eao-slam_b200/lib/variants/libsdm_warp_pieces.so; do
  name=$(basename $lib .so)
  SDM_LIB=$PWD/$lib timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2z_c${c}_${name}.json 2> gpurun_out/r2z_c${c}_${name}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2z_c${c}_${name}.json"))
print("config $c ${name}", round(d["ms_per_step"],3), round(d["kernel_ms_rank0"]["pass1_scan_ms"],3), d["fused_per_step_rank0"])
PY
done
done
timeout 900 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2z_tests.log 2>&1; echo tests rc=$?; tail -2 gpurun_out/r2z_tests.log
