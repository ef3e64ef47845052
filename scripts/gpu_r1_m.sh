#!/bin/bash
# Two-stage mode (coherent primary stage + megakernel on survivors) vs one-stage megakernel.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cp raytracer_go_b200/csrc/librt_b200.so gpurun_out/librt_b200_r1m.so
RT_B200_KERNEL=split timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_m_split.log 2>&1; echo "pytest split rc=$?" > gpurun_out/summary_m.txt
tail -8 gpurun_out/pytest_gpu_m_split.log
run() { label="$1"; shift
  env "$@" timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', round(d['value'],1),'Msamples/s', round(d['mrays_s'],1),'Mrays/s ms/step', round(d['ms_per_step'],2), 'share', r['kernel_share_of_step'])" >> gpurun_out/variants_m.txt 2>&1
}
run "mega" RT_B200_KERNEL=mega
run "split" RT_B200_KERNEL=split
run "split regen1" RT_B200_KERNEL=split RT_B200_REGEN_MIN=1
run "split regen16" RT_B200_KERNEL=split RT_B200_REGEN_MIN=16
run "mega again" RT_B200_KERNEL=mega
RT_B200_KERNEL=split timeout 200 python bench.py --config CB --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cornell split', round(d['value'],1), round(d['mrays_s'],1))" >> gpurun_out/variants_m.txt
timeout 200 python bench.py --config CB --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cornell mega', round(d['value'],1), round(d['mrays_s'],1))" >> gpurun_out/variants_m.txt
cat gpurun_out/variants_m.txt
CMD="python bench.py --spp 39 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
RT_B200_KERNEL=split $CMD > gpurun_out/plain_m.log 2>&1 && \
RT_B200_KERNEL=split ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:"render_kernel|primary_stage" -s 2 -c 2 --csv --log-file gpurun_out/split_kernels_m.csv $CMD > gpurun_out/ncu_m.log 2>&1
cat gpurun_out/split_kernels_m.csv | tail -10 | cut -c1-400
cat gpurun_out/summary_m.txt
