#!/bin/bash
# Warp-cooperative unit-sphere sampling (rand_unit_warp): parity first (under a short timeout), then speed.
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_aq.txt; : > $S
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_aq.log 2>&1; echo "smoke rc=$?" | tee -a $S; tail -2 gpurun_out/smoke_aq.log | tee -a $S
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_aq.txt 2>&1; echo "pytest rc=$?" | tee -a $S; tail -3 gpurun_out/pytest_aq.txt | tee -a $S
for cfg in C2 CB C4 C3; do
  timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms/step frac', round(d['roofline']['frac'],4))" | tee -a $S
done
