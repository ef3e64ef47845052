#!/bin/bash
# ncu --set full of one pass (72 spp) with the centre/half-extent slab test.
set -u
mkdir -p gpurun_out
CMD2="python bench.py --spp 72 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/plain_aa.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"render_kernel|primary_stage" -s 2 -c 2 -o gpurun_out/prof_r1aa $CMD2 > gpurun_out/ncu_full_aa.log 2>&1
echo "ncu rc=$?"
