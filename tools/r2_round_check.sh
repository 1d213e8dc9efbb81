#!/bin/bash
# round-2 final evidence set (single GPU), final build with the Edge Drawing kernels: suite, smoke, bench + reference arm, launch list, ncu --set full of the path's
# kernels and of the line-fit kernel.  Summaries are made from the .ncu-rep files afterwards (tools/ncu_summary.py, no GPU).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02o_suite.log 2>&1; echo suite rc=$?; tail -3 gpurun_out/r02o_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02o_smoke.log 2>&1; echo smoke rc=$?
python bench.py > gpurun_out/r02o_bench_c2_n1.json 2> gpurun_out/r02o_bench_c2_n1.err; echo bench rc=$?
python bench.py --impl reference > gpurun_out/r02o_ref.json 2> gpurun_out/r02o_ref.err; echo ref rc=$?
P="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-hot-spin"
$P > gpurun_out/r02o_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02o_launches.csv $P > gpurun_out/r02o_ncu_l.log 2>&1; echo ncu-l rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_lane|k_pass2_cand|k_intra_cand" -s 12 -c 4 -f -o gpurun_out/prof_r02o $P > gpurun_out/r02o_ncu_f.log 2>&1; echo ncu-f rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_pack|k_skip" -c 6 -f -o gpurun_out/prof_r02o_pack $P > gpurun_out/r02o_ncu_p.log 2>&1; echo ncu-p rc=$?
python tools/linefit_bench.py > gpurun_out/r02o_linefit_bench.json 2> gpurun_out/r02o_linefit_bench.err && \
ncu --set full --clock-control none --import-source on -k regex:k_line_fit -s 1 -c 1 -f -o gpurun_out/prof_r02o_linefit python tools/linefit_bench.py > gpurun_out/r02o_ncu_lf.log 2>&1; echo ncu-lf rc=$?
python tools/ed_bench.py --n 200 --out gpurun_out/r02o_ed_bench.json > gpurun_out/r02o_ed_bench.log 2>&1; echo ed-bench rc=$?
SDM_ED_ROUTE_PROF=1 python tools/ed_bench.py --n 32 --n-device 200 1000 > /dev/null 2> gpurun_out/r02o_ed_route_phases.log; grep k_ed_route gpurun_out/r02o_ed_route_phases.log | awk 'NR%3==0' > gpurun_out/r02o_ed_route_phases.txt
ncu --set full --clock-control none --import-source on -k regex:k_ed_route -s 1 -c 1 -f -o gpurun_out/prof_r02o_ed python tools/ed_bench.py --n 32 --n-device 200 > gpurun_out/r02o_ncu_ed.log 2>&1; echo ncu-ed rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02o_bench_c2_n1.json"))
print(d["ms_per_step"], d["value"], d["kernel_ms_rank0"], d["roofline"]["frac"], d["roofline"]["whole_path_frac"], d["clocks"], d["gpu_launches"])
for k in ("e2e","e2e_class","e2e_online","e2e_point_export","e2e_image_in_points_out","e2e_image_in_edge_drawing_points_out","edge_drawing"):
    if k in d: print(k, {a:b for a,b in d[k].items() if a not in ("api","note")})
print(d["cpu_baseline"])
PY
