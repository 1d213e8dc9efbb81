#!/bin/bash
# occupancy curve of the third-generation scan kernel on config 4 (long scans) and config 2, same box
mkdir -p gpurun_out
for c in 4 2; do
for lib in eao-slam_b200/lib/libsdm_b200.so eao-slam_b200/lib/variants/libsdm_minb8.so eao-slam_b200/lib/variants/libsdm_minb7.so eao-slam_b200/lib/variants/libsdm_minb6.so eao-slam_b200/lib/variants/libsdm_minb5.so eao-slam_b200/lib/variants/libsdm_minb4.so; do
  name=$(basename $lib .so)
  SDM_LIB=$PWD/$lib python bench.py --config $c --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2s_c${c}_${name}.json 2> gpurun_out/r2s_c${c}_${name}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2s_c${c}_${name}.json"))
print("config $c ${name}", round(d["ms_per_step"],3), round(d["kernel_ms_rank0"]["pass1_scan_ms"],3))
PY
done
done
