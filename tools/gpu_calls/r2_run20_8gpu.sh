#!/bin/bash
# 8 GPUs (charged 8x): the north-star line (config 3: 1000 keyframes over 8 GPUs) with cost-balanced and with equal-count shards
mkdir -p gpurun_out
python bench.py --gpus 8 --config 3 --no-e2e-variants --no-cpu-baseline > gpurun_out/r2x_c3_n8_balanced.json 2> gpurun_out/r2x_c3_n8_balanced.err; echo bal rc=$?; tail -2 gpurun_out/r2x_c3_n8_balanced.err
python bench.py --gpus 8 --config 3 --no-e2e --no-cpu-baseline --no-balance --no-parity > gpurun_out/r2x_c3_n8_equal.json 2> gpurun_out/r2x_c3_n8_equal.err; echo eq rc=$?
python - <<'PY'
import json
for n in ("balanced","equal"):
    try:
        d=json.load(open(f"gpurun_out/r2x_c3_n8_{n}.json"))
        print(n, round(d["ms_per_step"],3), d["kernel_ms_rank0"], "value %.3e"%d["value"], round(d["roofline"]["whole_path_frac"],4), d.get("sharding"), {k:d.get(k) for k in ("parity_checked_ranks","parity_boundary_keyframes","parity_boundary_mismatch_words")})
        if "e2e" in d: print("   e2e", {a:(round(b,3) if isinstance(b,float) else b) for a,b in d["e2e"].items() if a not in ("api","note")})
    except Exception as e:
        print(n, "failed", e)
PY
