#!/bin/bash
# after the robustness changes of the device routing (bounded gather, stale-plane reset, batch halving): ED + shim tests, smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edge_drawing.py tests/test_cpp_shim.py tests/test_abi.py -x -q > gpurun_out/r2ao_tests.log 2>&1; echo tests rc=$?
tail -2 gpurun_out/r2ao_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ao_smoke.log 2>&1; echo smoke rc=$?
