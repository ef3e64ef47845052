#!/bin/bash
# A/B: 512-thread CTAs (2 per SM, 32 warps at 64 registers) for the primary stage and / or the secondary
# megakernel against the default 256 x 3 (24 warps).
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_aj.txt; : > $S
run() { # label, env...
  local label=$1; shift
  for cfg in C2 CB; do
    env "$@" timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', '$cfg', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms/step')" | tee -a $S
  done
}
run default A=1
run pblock512 RT_B200_PBLOCK=512
run block512 RT_B200_BLOCK=512
run both512 RT_B200_PBLOCK=512 RT_B200_BLOCK=512
run default_again A=1
