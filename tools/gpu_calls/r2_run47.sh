#!/bin/bash
# longest paths of a tree by one sweep (no traversal per queried chain): parity tests, phase counts, timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_edge_drawing.py -x -q > gpurun_out/r2au_tests.log 2>&1; echo tests rc=$?
tail -1 gpurun_out/r2au_tests.log
SDM_ED_ROUTE_PROF=1 timeout 300 python tools/ed_bench.py --n 200 --n-device 200 1000 --out gpurun_out/r2au_ed_bench.json > gpurun_out/r2au_ed_bench.log 2>&1; echo bench rc=$?
grep k_ed_route gpurun_out/r2au_ed_bench.log | awk 'NR%3==0' | tail -2
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2au_ed_bench.json"))
for r in d["runs"]: print(r["threads"], round(r["wall_ms_per_kf"],4), round(r["route_thread_ms_per_kf"],3))
for r in d["device_route"]: print(r)
PY
