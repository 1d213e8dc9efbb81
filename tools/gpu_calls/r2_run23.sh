#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/linefit_bench.py > gpurun_out/r2ab_linefit_bench.json 2> gpurun_out/r2ab_linefit_bench.err; echo rc=$?; cat gpurun_out/r2ab_linefit_bench.json; tail -2 gpurun_out/r2ab_linefit_bench.err
