#!/bin/bash
# cache-policy variants of the skip walk's loads (ipair L1::no_allocate, skip byte L1::evict_last, texel L1::evict_first), configs 2 and 4
mkdir -p gpurun_out
for c in 2 4; do
for lib in eao-slam_b200/lib/libsdm_b200.so eao-slam_b200/lib/variants/libsdm_IPAIR_NOALLOC.so eao-slam_b200/lib/variants/libsdm_SKIP_EVICT_LAST.so eao-slam_b200/lib/variants/libsdm_NOALLOC_EVLAST.so eao-slam_b200/lib/variants/libsdm_TEXEL_EVICT_FIRST.so eao-slam_b200/lib/libsdm_b200.so; do
  name=$(basename $lib .so)
  SDM_LIB=$PWD/$lib python bench.py --config $c --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2u_c${c}_${name}.json 2> gpurun_out/r2u_c${c}_${name}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2u_c${c}_${name}.json"))
print("config $c ${name}", round(d["ms_per_step"],3), round(d["kernel_ms_rank0"]["pass1_scan_ms"],3), d["fused_per_step_rank0"])
PY
done
done
