#!/bin/bash
# Traversal stack in shared memory also when the scene stays in global memory (C4): parity, then speed.
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_as.txt; : > $S
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_as.txt 2>&1; echo "pytest rc=$?" | tee -a $S; tail -3 gpurun_out/pytest_as.txt | tee -a $S
for cfg in C4 C2; do
  timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms/step frac', round(d['roofline']['frac'],4))" | tee -a $S
done
RT_B200_NO_SMEM=1 timeout 300 python bench.py --config C2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 nosmem', round(d['value'],1), 'Msamples/s')" | tee -a $S
