#!/bin/bash
# 8 GPUs (charged 8x), final build: the north-star line (config 3: 1000 keyframes over 8 GPUs, cost-balanced shards, boundary parity)
mkdir -p gpurun_out
timeout 150 python bench.py --gpus 8 --config 3 --no-e2e --no-cpu-baseline > gpurun_out/r02t_c3_n8.json 2> gpurun_out/r02t_c3_n8.err; echo rc=$?; tail -2 gpurun_out/r02t_c3_n8.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02t_c3_n8.json"))
print(round(d["ms_per_step"],3), d["kernel_ms_rank0"], "value %.3e"%d["value"], round(d["roofline"]["whole_path_frac"],4), d.get("sharding"), {k:d.get(k) for k in ("parity_checked_ranks","parity_boundary_keyframes","parity_boundary_mismatch_words")})
PY
