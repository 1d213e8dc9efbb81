#!/bin/bash
# final bench line of the round (config 2, one GPU) with the detector-in-the-loop variant in both routing modes
mkdir -p gpurun_out
python bench.py > gpurun_out/r02o_bench_c2_n1.json 2> gpurun_out/r02o_bench_c2_n1.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02o_bench_c2_n1.json"))
print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["clocks"], d["gpu_launches"])
for k in ("e2e","e2e_image_in_points_out","e2e_image_in_edge_drawing_points_out","edge_drawing"):
    if k in d: print(k, {a:b for a,b in d[k].items() if a not in ("api","note")})
PY
