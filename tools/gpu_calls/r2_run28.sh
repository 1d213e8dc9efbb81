#!/bin/bash
# k_ed_planes4 (four pixels per thread, packed 16-bit lanes) + adaptive chunks: parity tests, timing tool, ncu --set full
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edge_drawing.py tests/test_cpp_shim.py -x -q -m gpu > gpurun_out/r2af_tests.log 2>&1; echo tests rc=$?
tail -5 gpurun_out/r2af_tests.log
timeout 600 python tools/ed_bench.py --n 200 --out gpurun_out/r2af_ed_bench.json > gpurun_out/r2af_ed_bench.log 2>&1; echo bench rc=$?
tail -2 gpurun_out/r2af_ed_bench.log | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ed_planes -s 30 -c 1 -f -o gpurun_out/prof_r02m_ed4 python tools/ed_bench.py --n 64 > gpurun_out/r2af_ncu.log 2>&1; echo ncu rc=$?
