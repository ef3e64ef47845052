#!/bin/bash
# A/B inside one call (same box): balanced vs greedy passes, pass size; then ncu --set full of one balanced pass.
set -u
mkdir -p gpurun_out; rm -f gpurun_out/variants_y.txt
run() { label="$1"; shift
  env "$@" timeout 200 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', round(d['value'],1),'Msamples/s ms/step', round(d['ms_per_step'],2), 'ms/launch', round(r['ms_per_launch'],3), 'launches', r['launches'])" >> gpurun_out/variants_y.txt 2>&1
}
run "balanced 64M" A=1
run "greedy 64M" RT_B200_PASS_BALANCE=0
run "balanced 64M again" A=1
run "greedy 64M again" RT_B200_PASS_BALANCE=0
run "balanced 128M" RT_B200_PASS_PATHS=134217728
run "greedy 128M" RT_B200_PASS_PATHS=134217728 RT_B200_PASS_BALANCE=0
run "balanced 256M" RT_B200_PASS_PATHS=268435456
run "balanced 32M" RT_B200_PASS_PATHS=33554432
cat gpurun_out/variants_y.txt
CMD2="python bench.py --spp 72 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/plain_y2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"render_kernel|primary_stage" -s 2 -c 2 -o gpurun_out/prof_r1y $CMD2 > gpurun_out/ncu_full_y.log 2>&1
echo "ncu rc=$?"
