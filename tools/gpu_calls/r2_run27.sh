#!/bin/bash
# Edge Drawing on the device after the odd-width alignment fix: parity tests (C-ABI + class shim), timing tool, ncu --set full of k_ed_planes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edge_drawing.py tests/test_cpp_shim.py -x -q -m gpu > gpurun_out/r2ae_tests.log 2>&1; echo tests rc=$?
tail -5 gpurun_out/r2ae_tests.log
timeout 600 python tools/ed_bench.py --n 200 --out gpurun_out/r2ae_ed_bench.json > gpurun_out/r2ae_ed_bench.log 2>&1; echo bench rc=$?
tail -2 gpurun_out/r2ae_ed_bench.log | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ed_planes -s 30 -c 1 -f -o gpurun_out/prof_r02l_ed python tools/ed_bench.py --n 64 > gpurun_out/r2ae_ncu.log 2>&1; echo ncu rc=$?
