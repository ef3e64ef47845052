#!/bin/bash
# ncu full capture of one pass of the final round-1 build
set -u
mkdir -p gpurun_out
CMD2="python bench.py --spp 82 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/plain_bc.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"render_kernel|primary_stage" -s 2 -c 2 -o gpurun_out/prof_r1bc $CMD2 > gpurun_out/ncu_full_bc.log 2>&1
echo "ncu rc=$?"
