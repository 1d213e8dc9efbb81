#!/bin/bash
# round 2, 2-GPU call: multi-rank parity on two real devices (NVLink peers), device-ordered vs host-ordered exchange
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1
python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/r2g_multirank.log 2>&1; echo multirank rc=$?; tail -3 gpurun_out/r2g_multirank.log
python bench.py --gpus 2 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo bench-n2 rc=$?; tail -2 gpurun_out/r2g_bench_n2.err
SDM_BENCH_HOST_EXCHANGE=1 python bench.py --gpus 2 --no-e2e --no-parity > gpurun_out/r2g_bench_n2_hostx.json 2> gpurun_out/r2g_bench_n2_hostx.err; echo bench-n2-hostx rc=$?
python bench.py --gpus 2 --no-e2e --no-parity > gpurun_out/r2g_bench_n2_devx.json 2> gpurun_out/r2g_bench_n2_devx.err; echo bench-n2-devx rc=$?
python bench.py --gpus 1 --no-e2e --no-cpu-baseline > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo bench-n1 rc=$?
python bench.py --gpus 2 --config 3 --no-e2e-variants > gpurun_out/r2g_bench_c3_n2.json 2> gpurun_out/r2g_bench_c3_n2.err; echo bench-c3-n2 rc=$?
python bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r2g_ref_n2.json 2> gpurun_out/r2g_ref_n2.err; echo ref-n2 rc=$?
python - <<'PY'
import json
for n in ("n2","n2_hostx","n2_devx","n1","c3_n2"):
    try:
        d=json.load(open(f"gpurun_out/r2g_bench_{n}.json"))
        print(n, round(d["ms_per_step"],3), d["kernel_ms_rank0"], "value %.3e"%d["value"], {k:d.get(k) for k in ("parity_checked_ranks","parity_boundary_keyframes","parity_boundary_mismatch_words","exchange")})
        for k in ("e2e","e2e_dense"):
            if k in d: print("   ", k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in d[k].items() if a not in ("api","note","sample")})
    except Exception as e:
        print(n, "failed", e)
d=json.load(open("gpurun_out/r2g_ref_n2.json")); print("ref arm cores", d["cpu_baseline"]["cores"], d["value"])
PY
