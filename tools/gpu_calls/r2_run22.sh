#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linefit.py -x -q -s > gpurun_out/r2aa_linefit.log 2>&1; echo linefit rc=$?; tail -12 gpurun_out/r2aa_linefit.log
