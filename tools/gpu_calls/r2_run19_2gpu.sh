#!/bin/bash
# 2 GPUs: balanced (equal estimated cost) vs equal-count shards of config 3, boundary parity on both; GPU multirank test
mkdir -p gpurun_out
python bench.py --gpus 2 --config 3 --no-e2e --no-cpu-baseline > gpurun_out/r2w_c3_n2_balanced.json 2> gpurun_out/r2w_c3_n2_balanced.err; echo bal rc=$?; tail -2 gpurun_out/r2w_c3_n2_balanced.err
python bench.py --gpus 2 --config 3 --no-e2e --no-cpu-baseline --no-balance > gpurun_out/r2w_c3_n2_equal.json 2> gpurun_out/r2w_c3_n2_equal.err; echo eq rc=$?
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/r2w_multirank.log 2>&1; echo multirank rc=$?; tail -2 gpurun_out/r2w_multirank.log
python - <<'PY'
import json
for n in ("balanced","equal"):
    d=json.load(open(f"gpurun_out/r2w_c3_n2_{n}.json"))
    print(n, round(d["ms_per_step"],3), d["kernel_ms_rank0"], d.get("sharding"), {k:d.get(k) for k in ("parity_checked_ranks","parity_boundary_keyframes","parity_boundary_mismatch_words")})
PY
