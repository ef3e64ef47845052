#!/bin/bash
# Knob sweep under the 512 x 2 default (regeneration threshold, pass size, chunk, stages), then the ncu
# full capture of one pass of both kernels and the launch list of the bench command.
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_am.txt; : > $S
run() { local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms/step')" | tee -a $S; }
run default A=1
run regen4 RT_B200_REGEN_MIN=4
run regen8 RT_B200_REGEN_MIN=8
run regen16 RT_B200_REGEN_MIN=16
run pass128M RT_B200_PASS_PATHS=134217728
run pass32M RT_B200_PASS_PATHS=33554432
run chunk128 RT_B200_CHUNK=128
run chunk1024 RT_B200_CHUNK=1024
run stages2 RT_B200_STAGES=2
run leaf2 RT_B200_MAX_LEAF=2
run leaf8 RT_B200_MAX_LEAF=8
run default_again A=1
CMD2="python bench.py --spp 82 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/plain_am.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"render_kernel|primary_stage" -s 2 -c 2 -o gpurun_out/prof_r1am $CMD2 > gpurun_out/ncu_full_am.log 2>&1
echo "ncu rc=$?" | tee -a $S
