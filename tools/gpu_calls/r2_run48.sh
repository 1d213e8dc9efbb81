#!/bin/bash
# last call of the round: the final library through the class shim's Edge Drawing tests and smoke()
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_cpp_shim.py -x -q -k "edge" > gpurun_out/r02u_shim_tests.log 2>&1; echo tests rc=$?; tail -1 gpurun_out/r02u_shim_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02u_smoke.log 2>&1; echo smoke rc=$?
