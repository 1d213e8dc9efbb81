#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2h_suite.log 2>&1; echo suite rc=$?; tail -5 gpurun_out/r2h_suite.log
python bench.py > gpurun_out/r2h_bench_c2.json 2> gpurun_out/r2h_bench_c2.err; echo bench-c2 rc=$?; tail -3 gpurun_out/r2h_bench_c2.err
python bench.py --config 4 --no-e2e-variants > gpurun_out/r2h_bench_c4.json 2> gpurun_out/r2h_bench_c4.err; echo bench-c4 rc=$?; tail -3 gpurun_out/r2h_bench_c4.err
SDM_SCAN=lane2 python bench.py --config 4 --no-e2e --no-cpu-baseline > gpurun_out/r2h_bench_c4_lane2.json 2> gpurun_out/r2h_bench_c4_lane2.err; echo bench-c4-lane2 rc=$?
for g in lane3 lane2; do
SDM_SCAN=$g ncu --set full --clock-control none --import-source on -k regex:"k_pass1_lane" -s 3 -c 1 -f -o gpurun_out/prof_r2h_c4_$g python bench.py --config 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-hot-spin > gpurun_out/ncu_r2h_c4_$g.log 2>&1; echo ncu-c4-$g rc=$?
done
python - <<'PY'
import json
for n in ("c2","c4","c4_lane2"):
    try:
        d=json.load(open(f"gpurun_out/r2h_bench_{n}.json"))
        print(n, round(d["ms_per_step"],3), d["kernel_ms_rank0"], "value %.3e"%d["value"], "cands", d["candidates_per_step"], "frac", round(d["roofline"]["frac"],4), round(d["roofline"]["whole_path_frac"],4), round(d["roofline"]["whole_path_frac_with_compaction"],4), d.get("scan_generation"))
        for k in ("e2e","e2e_dense","e2e_scatter","e2e_point_export","e2e_image_in_points_out","e2e_class","cpu_baseline"):
            if k in d: print("   ", k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in d[k].items() if a not in ("api","note","sample")})
    except Exception as e:
        print(n, "failed", e)
PY
