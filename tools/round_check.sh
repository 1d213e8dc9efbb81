set -x
python -m pytest tests -x -q -m gpu > gpurun_out/s31_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/s31_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s31_smoke.log 2>&1; echo smoke rc=$?
python bench.py > gpurun_out/s31_bench_n1.json 2> gpurun_out/s31_bench_n1.err; echo bench rc=$?
python bench.py --impl reference > gpurun_out/s31_ref.json 2> gpurun_out/s31_ref.err; echo ref rc=$?
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-hot-spin > gpurun_out/plain_r01i.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01i.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-hot-spin > gpurun_out/ncu_l_r01i.log 2>&1; echo ncu-l rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_lane|k_pass2_cand|k_intra_cand" -s 12 -c 4 -f -o gpurun_out/prof_r01i python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-hot-spin > gpurun_out/ncu_f_r01i.log 2>&1; echo ncu-f rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/s31_bench_n1.json"))
print(d["ms_per_step"], d["value"], d["kernel_ms_rank0"], d["roofline"]["frac"], d["roofline"]["whole_path_frac"], d["clocks"], d["gpu_launches"])
for k in ("e2e","e2e_scatter","e2e_point_export","e2e_image_in_points_out"):
    print(k, {a:b for a,b in d[k].items() if a not in ("api","note")})
print(d["cpu_baseline"])
PY
