#!/bin/bash
# round 2, 8-GPU call: BASELINE configs 2 (weak, 200 keyframes per GPU), 3 (1000 keyframes / 8 GPUs: the north-star line) and
# 5 (4096 keyframes / 8 GPUs), reference arm under torchrun
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2j_topo8.txt 2>&1
python bench.py --gpus 8 --config 3 --no-e2e-variants > gpurun_out/r2j_bench_c3_n8.json 2> gpurun_out/r2j_bench_c3_n8.err; echo c3-n8 rc=$?; tail -2 gpurun_out/r2j_bench_c3_n8.err
python bench.py --gpus 8 > gpurun_out/r2j_bench_c2_n8.json 2> gpurun_out/r2j_bench_c2_n8.err; echo c2-n8 rc=$?; tail -2 gpurun_out/r2j_bench_c2_n8.err
python bench.py --gpus 8 --config 5 --no-parity > gpurun_out/r2j_bench_c5_n8.json 2> gpurun_out/r2j_bench_c5_n8.err; echo c5-n8 rc=$?; tail -2 gpurun_out/r2j_bench_c5_n8.err
python bench.py --gpus 8 --impl reference --steps 2 --warmup 1 > gpurun_out/r2j_ref_n8.json 2> gpurun_out/r2j_ref_n8.err; echo ref-n8 rc=$?
python - <<'PY'
import json
for n in ("c3_n8","c2_n8","c5_n8"):
    try:
        d=json.load(open(f"gpurun_out/r2j_bench_{n}.json"))
        print(n, round(d["ms_per_step"],3), d["kernel_ms_rank0"], "value %.3e"%d["value"], "whole", round(d["roofline"]["whole_path_frac"],4), {k:d.get(k) for k in ("parity_checked_ranks","parity_boundary_keyframes","parity_boundary_mismatch_words")})
        for k in ("e2e","e2e_point_export","e2e_image_in_points_out","north_star"):
            if k in d: print("   ", k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in d[k].items() if a not in ("api","note","sample")})
    except Exception as e:
        print(n, "failed", e)
d=json.load(open("gpurun_out/r2j_ref_n8.json")); print("ref arm cores", d["cpu_baseline"]["cores"], d["value"])
PY
