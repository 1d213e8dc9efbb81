#!/bin/bash
# 4 GPUs: the missing point of the strong-scaling curve of config 3 (1 / 2 / 4 / 8), cost-balanced shards
mkdir -p gpurun_out
python bench.py --gpus 4 --config 3 --no-e2e --no-cpu-baseline > gpurun_out/r2ad_c3_n4.json 2> gpurun_out/r2ad_c3_n4.err; echo rc=$?; tail -2 gpurun_out/r2ad_c3_n4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2ad_c3_n4.json"))
print(round(d["ms_per_step"],3), d["kernel_ms_rank0"], "value %.3e"%d["value"], round(d["roofline"]["whole_path_frac"],4), d.get("sharding",{}).get("bounds"), {k:d.get(k) for k in ("parity_checked_ranks","parity_boundary_keyframes","parity_boundary_mismatch_words")})
PY
