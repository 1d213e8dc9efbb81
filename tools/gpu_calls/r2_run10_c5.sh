#!/bin/bash
# configs 3 and 5 at N GPUs (N = the GPUs of this call): bash tools/r2_run10_c5.sh N
N=$1
mkdir -p gpurun_out
python bench.py --gpus $N --config 5 --no-parity > gpurun_out/r2k_bench_c5_n$N.json 2> gpurun_out/r2k_bench_c5_n$N.err; echo c5-n$N rc=$?; tail -2 gpurun_out/r2k_bench_c5_n$N.err
python bench.py --gpus $N --config 3 --no-e2e-variants > gpurun_out/r2k_bench_c3_n$N.json 2> gpurun_out/r2k_bench_c3_n$N.err; echo c3-n$N rc=$?; tail -2 gpurun_out/r2k_bench_c3_n$N.err
if [ "$N" = "1" ]; then
  python bench.py --config 5 --impl reference --steps 1 --warmup 0 > gpurun_out/r2k_ref_c5.json 2> gpurun_out/r2k_ref_c5.err; echo ref-c5 rc=$?
  python bench.py --config 4 --impl reference --steps 1 --warmup 0 > gpurun_out/r2k_ref_c4.json 2> gpurun_out/r2k_ref_c4.err; echo ref-c4 rc=$?
fi
python - <<PY
import json
for n in ("c5_n$N","c3_n$N"):
    try:
        d=json.load(open(f"gpurun_out/r2k_bench_{n}.json"))
        print(n, round(d["ms_per_step"],3), d["kernel_ms_rank0"], "value %.3e"%d["value"], "whole", round(d["roofline"]["whole_path_frac"],4), {k:d.get(k) for k in ("parity_checked_ranks","parity_boundary_mismatch_words")})
        for k in ("e2e","cpu_baseline"):
            if k in d: print("   ", k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in d[k].items() if a not in ("api","note","sample")})
    except Exception as e:
        print(n, "failed", e)
PY
