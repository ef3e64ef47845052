#!/bin/bash
# One queue-append atomic per CTA and round in the primary stage (instead of one per warp): parity, then speed.
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_av.txt; : > $S
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_av.txt 2>&1; echo "pytest rc=$?" | tee -a $S; tail -3 gpurun_out/pytest_av.txt | tee -a $S
for cfg in C2 CB C4 C3 C1; do
  timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms/step frac', round(d['roofline']['frac'],4))" | tee -a $S
done
