#!/bin/bash
# k_ed_route: split of the segment extraction (SDM_ED_ROUTE_PROF)
mkdir -p gpurun_out
SDM_ED_ROUTE_PROF=1 timeout 300 python tools/ed_bench.py --n 8 --n-device 200 > gpurun_out/r2at_ed_bench.log 2>&1; echo bench rc=$?
grep k_ed_route gpurun_out/r2at_ed_bench.log | tail -1
