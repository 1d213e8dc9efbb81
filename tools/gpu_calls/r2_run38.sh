#!/bin/bash
# leaner walk step (pointers, 32-bit offsets, one-level exits): parity tests, per-phase cycle counts
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edge_drawing.py -x -q > gpurun_out/r2ap_tests.log 2>&1; echo tests rc=$?
tail -2 gpurun_out/r2ap_tests.log
SDM_ED_ROUTE_PROF=1 timeout 900 python tools/ed_bench.py --n 32 --n-device 200 1000 > gpurun_out/r2ap_ed_bench.log 2>&1; echo bench rc=$?
grep k_ed_route gpurun_out/r2ap_ed_bench.log | awk 'NR%3==0' | tail -3
