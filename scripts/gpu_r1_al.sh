#!/bin/bash
# A/B: secondary megakernel at 640 x 2 (40 warps, 48 registers, some spills) against 512 x 2.
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_al.txt; : > $S
run() { local label=$1; shift
  for cfg in C2 CB C4; do
    env "$@" timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', '$cfg', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms/step')" | tee -a $S
  done; }
run b512 A=1
run b640 RT_B200_BLOCK=640
