#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linefit.py -x -q -s > gpurun_out/r2p_linefit.log 2>&1; echo linefit rc=$?; tail -8 gpurun_out/r2p_linefit.log
timeout 600 python tools/linefit_bench.py > gpurun_out/r2p_linefit_bench.json 2> gpurun_out/r2p_linefit_bench.err; echo lfbench rc=$?; cat gpurun_out/r2p_linefit_bench.json; tail -3 gpurun_out/r2p_linefit_bench.err
