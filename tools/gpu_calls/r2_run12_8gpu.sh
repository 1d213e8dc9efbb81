#!/bin/bash
# round 2, 8-GPU box (charged 8x: keep it short): BASELINE config 3 at N=8 (1000 keyframes over 8 GPUs: the north-star line) and
# the config-5 sweep points N=8 / 4 / 2 (4096 keyframes; N=1 and the 64-keyframe CPU arm come from the single-GPU call r2k)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m_topo8.txt 2>&1
python bench.py --gpus 8 --config 3 --no-e2e-variants > gpurun_out/r2m_bench_c3_n8.json 2> gpurun_out/r2m_bench_c3_n8.err; echo c3-n8 rc=$?; tail -2 gpurun_out/r2m_bench_c3_n8.err
for n in 8 4 2; do
  python bench.py --gpus $n --config 5 --no-cpu-baseline --no-e2e > gpurun_out/r2m_bench_c5_n$n.json 2> gpurun_out/r2m_bench_c5_n$n.err; echo c5-n$n rc=$?; tail -2 gpurun_out/r2m_bench_c5_n$n.err
done
python - <<'PY'
import json
for n in ("c3_n8","c5_n8","c5_n4","c5_n2"):
    try:
        d=json.load(open(f"gpurun_out/r2m_bench_{n}.json"))
        print(n, round(d["ms_per_step"],3), d["kernel_ms_rank0"], "value %.3e"%d["value"], "whole", round(d["roofline"]["whole_path_frac"],4), {k:d.get(k) for k in ("parity_checked_ranks","parity_boundary_keyframes","parity_boundary_mismatch_words")})
        for k in ("e2e","north_star"):
            if k in d: print("   ", k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in d[k].items() if a not in ("api","note","sample")})
    except Exception as e:
        print(n, "failed", e)
PY
