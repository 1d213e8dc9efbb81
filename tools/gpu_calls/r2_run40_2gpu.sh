#!/bin/bash
# final build on two GPUs: the driver's launch of bench.py at N = 2 (weak scaling, config 2), boundary parity on both ranks
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02o_bench_c2_n2.json 2> gpurun_out/r02o_bench_c2_n2.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02o_bench_c2_n2.json"))
print(d["n_gpus"], d["ms_per_step"], d["value"], d.get("parity"), d.get("parity_checked_ranks"), d["e2e"]["ms_per_step"] if "e2e" in d else None)
PY
tail -3 gpurun_out/r02o_bench_c2_n2.err
