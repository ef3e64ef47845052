#!/bin/bash
# Staged mode: how many coherent stage kernels before the megakernel?
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cp raytracer_go_b200/csrc/librt_b200.so gpurun_out/librt_b200_r1o.so
RT_B200_STAGES=3 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_o_s3.log 2>&1; echo "pytest stages3 rc=$?" > gpurun_out/summary_o.txt
tail -5 gpurun_out/pytest_gpu_o_s3.log
run() { label="$1"; shift
  env "$@" timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', round(d['value'],1),'Msamples/s', round(d['mrays_s'],1),'Mrays/s ms/step', round(d['ms_per_step'],2))" >> gpurun_out/variants_o.txt 2>&1
}
for st in 1 2 3 4 6 8; do run "stages $st" RT_B200_STAGES=$st; done
run "stages 2 regen8" RT_B200_STAGES=2 RT_B200_REGEN_MIN=8
for st in 1 2 3 4; do
RT_B200_STAGES=$st timeout 200 python bench.py --config CB --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cornell stages $st', round(d['value'],1), round(d['mrays_s'],1))" >> gpurun_out/variants_o.txt
done
for st in 1 2 3; do
RT_B200_STAGES=$st timeout 300 python bench.py --config C4 --spp 16 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 stages $st', round(d['value'],1), round(d['mrays_s'],1))" >> gpurun_out/variants_o.txt
done
cat gpurun_out/variants_o.txt
CMD="python bench.py --spp 39 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
RT_B200_STAGES=3 $CMD > gpurun_out/plain_o.log 2>&1 && \
RT_B200_STAGES=3 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"render_kernel|primary_stage" -s 4 -c 4 --csv --log-file gpurun_out/stages3_kernels_o.csv $CMD > gpurun_out/ncu_o.log 2>&1
grep -E "duration|thread_inst" gpurun_out/stages3_kernels_o.csv | cut -d, -f5,13,15
cat gpurun_out/summary_o.txt
