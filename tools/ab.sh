#!/bin/bash
# A/B timing of library variants on the GPU box: tools/ab.sh tag lib1 lib2 ...   (quick bench, resident loop only)
tag=$1; shift
for lib in "$@"; do
  name=$(basename $lib .so)
  SDM_LIB=$PWD/$lib python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_${name}.json"))
print("${name}", round(d["ms_per_step"],3), d["kernel_ms_rank0"], d.get("fused_per_step_rank0"), d.get("checked_per_step_rank0"))
PY
done
