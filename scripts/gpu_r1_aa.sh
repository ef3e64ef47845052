#!/bin/bash
# ncu --set full of one pass (82 spp = the greedy pass size of C2) of the current kernels.
set -u
mkdir -p gpurun_out
CMD2="python bench.py --spp 82 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/plain_af.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"render_kernel|primary_stage" -s 2 -c 2 -o gpurun_out/prof_r1af $CMD2 > gpurun_out/ncu_full_af.log 2>&1
echo "ncu rc=$?"
