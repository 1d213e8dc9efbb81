#!/bin/bash
# k_ed_sort (eight warps per image) feeds both routing modes: parity tests, timing tool (host threads + device), phase counts
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edge_drawing.py tests/test_cpp_shim.py -x -q > gpurun_out/r2as_tests.log 2>&1; echo tests rc=$?
tail -2 gpurun_out/r2as_tests.log
SDM_ED_ROUTE_PROF=1 timeout 900 python tools/ed_bench.py --n 200 --out gpurun_out/r2as_ed_bench.json > gpurun_out/r2as_ed_bench.log 2>&1; echo bench rc=$?
grep k_ed_route gpurun_out/r2as_ed_bench.log | awk 'NR%3==0' | tail -2
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2as_ed_bench.json"))
for r in d["runs"]: print(r["threads"], round(r["wall_ms_per_kf"],4), round(r["kernel_us_per_kf"],2), round(r["route_thread_ms_per_kf"],3))
for r in d["device_route"]: print(r)
print(d["host_only_one_thread_ms_per_kf"])
PY
