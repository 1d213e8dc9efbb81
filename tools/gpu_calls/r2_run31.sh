#!/bin/bash
# Edge Drawing stage 2 on the device, third form (anchors sorted by the whole warp): parity tests, timing tool
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edge_drawing.py tests/test_cpp_shim.py -x -q -m gpu -s > gpurun_out/r2aj_tests.log 2>&1; echo tests rc=$?
tail -5 gpurun_out/r2aj_tests.log
timeout 900 python tools/ed_bench.py --n 200 --out gpurun_out/r2aj_ed_bench.json > gpurun_out/r2aj_ed_bench.log 2>&1; echo bench rc=$?
tail -3 gpurun_out/r2aj_ed_bench.log | cut -c1-300
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2aj_ed_bench.json"))
for r in d["runs"]: print(r["threads"], round(r["wall_ms_per_kf"],4), round(r["kernel_us_per_kf"],2), round(r["route_thread_ms_per_kf"],3))
for r in d["device_route"]: print(r)
PY
