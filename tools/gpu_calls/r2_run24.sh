#!/bin/bash
# config 4 with the long-scan build: ncu --set full of the scan kernel (the r02b capture is of the 48-register build)
mkdir -p gpurun_out
P="python bench.py --config 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-hot-spin"
$P > gpurun_out/r2ac_c4_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pass1_lane -s 3 -c 1 -f -o gpurun_out/prof_r02k_c4 $P > gpurun_out/r2ac_ncu_c4.log 2>&1; echo ncu rc=$?
tail -1 gpurun_out/r2ac_c4_plain.log | cut -c1-300
