#!/bin/bash
# single GPU: suite (TMA variant under timeout first), config 3 at N=1 (1000 keyframes) with the step debug print, TMA A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_edge_cases.py -x -q -m gpu -k generations > gpurun_out/r2i_gen.log 2>&1; echo gen rc=$?; tail -3 gpurun_out/r2i_gen.log
python -m pytest tests -x -q -m gpu > gpurun_out/r2i_suite.log 2>&1; echo suite rc=$?; tail -4 gpurun_out/r2i_suite.log
SDM_BENCH_DEBUG=1 python bench.py --config 3 --steps 3 --no-e2e --no-cpu-baseline > gpurun_out/r2i_bench_c3_n1.json 2> gpurun_out/r2i_bench_c3_n1.err; echo c3 rc=$?; grep "step:" gpurun_out/r2i_bench_c3_n1.err | tail -4
SDM_BENCH_DEBUG=1 python bench.py --kf 500 --steps 3 --no-e2e --no-cpu-baseline > gpurun_out/r2i_bench_kf500.json 2> gpurun_out/r2i_bench_kf500.err; echo kf500 rc=$?; grep "step:" gpurun_out/r2i_bench_kf500.err | tail -3
for v in warp warp_tma; do
  SDM_SCAN=$v timeout 600 python bench.py --config 4 --kf 16 --steps 3 --no-e2e --no-cpu-baseline > gpurun_out/r2i_c4_$v.json 2> gpurun_out/r2i_c4_$v.err; echo c4-$v rc=$?
  SDM_SCAN=$v timeout 600 python bench.py --kf 40 --steps 3 --no-e2e --no-cpu-baseline > gpurun_out/r2i_c2_$v.json 2> gpurun_out/r2i_c2_$v.err; echo c2-$v rc=$?
done
SDM_SCAN=lane3 python bench.py --config 4 --kf 16 --steps 3 --no-e2e --no-cpu-baseline > gpurun_out/r2i_c4_lane3.json 2> gpurun_out/r2i_c4_lane3.err
python bench.py > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err; echo bench-c2 rc=$?
python - <<'PY'
import json
for n in ("bench_c3_n1","bench_kf500","c4_warp","c4_warp_tma","c4_lane3","c2_warp","c2_warp_tma","bench_c2"):
    try:
        d=json.load(open(f"gpurun_out/r2i_{n}.json"))
        print(n, round(d["ms_per_step"],3), d["kernel_ms_rank0"], "value %.3e"%d["value"], d.get("scan_generation"), d["fused_per_step_rank0"])
        for k in ("e2e","e2e_blocks","e2e_class","e2e_point_export","e2e_image_in_points_out"):
            if k in d: print("   ", k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in d[k].items() if a not in ("api","note","sample")})
    except Exception as e:
        print(n, "failed", e)
PY
