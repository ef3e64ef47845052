#!/bin/bash
# 512 x 2 CTAs as the default: tests in both kernel modes, every config against the previous 256 x 3.
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_ak.txt; : > $S
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ak.txt 2>&1; echo "pytest rc=$?" | tee -a $S; tail -1 gpurun_out/pytest_ak.txt | tee -a $S
RT_B200_KERNEL=mega timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ak_mega.txt 2>&1; echo "pytest mega rc=$?" | tee -a $S; tail -1 gpurun_out/pytest_ak_mega.txt | tee -a $S
run() { # label, env...
  local label=$1; shift
  for cfg in C1 C2 C3 C4 CB; do
    env "$@" timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', '$cfg', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms/step, e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4))" | tee -a $S
  done
}
run new_512x2 A=1
run old_256x3 RT_B200_BLOCK=256 RT_B200_PBLOCK=256
RT_B200_KERNEL=mega timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mega 512x2 C2', round(d['value'],1))" | tee -a $S
RT_B200_KERNEL=mega RT_B200_BLOCK=256 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mega 256x3 C2', round(d['value'],1))" | tee -a $S
