#!/bin/bash
# tree-local lists back in global memory (the shared-memory form cost instructions and gained nothing): parity tests, per-phase cycle counts
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edge_drawing.py -x -q > gpurun_out/r2aq_tests.log 2>&1; echo tests rc=$?
tail -2 gpurun_out/r2aq_tests.log
SDM_ED_ROUTE_PROF=1 timeout 900 python tools/ed_bench.py --n 32 --n-device 200 1000 > gpurun_out/r2aq_ed_bench.log 2>&1; echo bench rc=$?
grep k_ed_route gpurun_out/r2aq_ed_bench.log | awk 'NR%3==0' | tail -3
