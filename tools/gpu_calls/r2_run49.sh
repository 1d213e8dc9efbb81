#!/bin/bash
# SDM_ED_ROUTE_HOST_MASKS_ON_DEVICE (host walks, k_ed_mask): Edge Drawing tests, bench line with the three detector-in-the-loop modes
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_edge_drawing.py -x -q > gpurun_out/r02v_ed_tests.log 2>&1; echo tests rc=$?; tail -1 gpurun_out/r02v_ed_tests.log
timeout 100 python bench.py > gpurun_out/r02v_bench_c2_n1.json 2> gpurun_out/r02v_bench_c2_n1.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02v_bench_c2_n1.json"))
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])
k="e2e_image_in_edge_drawing_points_out"
print({a:b for a,b in d[k].items() if a not in ("api","note")})
PY
