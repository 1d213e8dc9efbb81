#!/bin/bash
# config 4 (long scans, latency-bound): occupancy / unroll variants of the third-generation scan kernel, same box
mkdir -p gpurun_out
for lib in eao-slam_b200/lib/libsdm_b200.so eao-slam_b200/lib/variants/libsdm_minb12.so eao-slam_b200/lib/variants/libsdm_minb8.so eao-slam_b200/lib/variants/libsdm_unroll1.so eao-slam_b200/lib/libsdm_b200.so; do
  name=$(basename $lib .so)
  SDM_LIB=$PWD/$lib python bench.py --config 4 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2r_c4_${name}.json 2> gpurun_out/r2r_c4_${name}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2r_c4_${name}.json"))
print("${name}", round(d["ms_per_step"],3), d["kernel_ms_rank0"])
PY
done
