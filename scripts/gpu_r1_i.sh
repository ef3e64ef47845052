#!/bin/bash
# Eighth GPU pass: full gpu suite, launch list and full ncu capture of the shipped kernels.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cp raytracer_go_b200/csrc/librt_b200.so gpurun_out/librt_b200_r1i.so
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary_i.txt
tail -4 gpurun_out/pytest_gpu_i.log
python __graft_entry__.py smoke > gpurun_out/smoke_i.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary_i.txt; tail -1 gpurun_out/smoke_i.log
# launch list of the default bench command (short steps)
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_i.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1i.csv $CMD > gpurun_out/ncu_launch_i.log 2>&1
# full capture of the megakernel at the real pass size (39 spp/pass x 810000 px = one 31.6 M-path launch)
CMD2="python bench.py --spp 39 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/plain_i2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/prof_r1i $CMD2 > gpurun_out/ncu_full_i.log 2>&1
CMD3="python bench.py --config CB --spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD3 > gpurun_out/plain_i3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/prof_r1i_cornell $CMD3 > gpurun_out/ncu_full_i3.log 2>&1
cat gpurun_out/summary_i.txt
