#!/bin/bash
# last build of the round (k_ed_sort): suite, smoke, bench + reference arm, Edge Drawing timing, ncu of k_ed_sort + k_ed_route
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02t_suite.log 2>&1; echo suite rc=$?; tail -2 gpurun_out/r02t_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02t_smoke.log 2>&1; echo smoke rc=$?
python bench.py > gpurun_out/r02t_bench_c2_n1.json 2> gpurun_out/r02t_bench_c2_n1.err; echo bench rc=$?
python bench.py --impl reference > gpurun_out/r02t_ref.json 2> gpurun_out/r02t_ref.err; echo ref rc=$?
python tools/ed_bench.py --n 200 --out gpurun_out/r02t_ed_bench.json > gpurun_out/r02t_ed_bench.log 2>&1; echo ed-bench rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_ed_sort|k_ed_route" -s 2 -c 2 -f -o gpurun_out/prof_r02t_ed_route python tools/ed_bench.py --n 8 --n-device 200 > gpurun_out/r02t_ncu_ed.log 2>&1; echo ncu-route rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02t_bench_c2_n1.json"))
print(d["ms_per_step"], d["value"], d["kernel_ms_rank0"], d["roofline"]["frac"], d["clocks"], d["gpu_launches"])
for k in ("e2e","e2e_class","e2e_image_in_points_out","e2e_image_in_edge_drawing_points_out","edge_drawing"):
    if k in d: print(k, {a:b for a,b in d[k].items() if a not in ("api","note")})
PY
