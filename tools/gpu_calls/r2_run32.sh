#!/bin/bash
# k_ed_route: per-phase cycle counts (SDM_ED_ROUTE_PROF) on 200 and 1000 keyframes
mkdir -p gpurun_out
SDM_ED_ROUTE_PROF=1 timeout 900 python tools/ed_bench.py --n 32 --out gpurun_out/r2ak_ed_bench.json > gpurun_out/r2ak_ed_bench.log 2>&1; echo bench rc=$?
grep k_ed_route gpurun_out/r2ak_ed_bench.log | tail -8
