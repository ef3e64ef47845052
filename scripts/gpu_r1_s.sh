#!/bin/bash
# New tests + ncu evidence for the BVH-fetch regime (C4, 1M spheres in global memory).
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cp raytracer_go_b200/csrc/librt_b200.so gpurun_out/librt_b200_r1s.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary_s.txt
tail -3 gpurun_out/pytest_gpu_s.log
CMD="python bench.py --config C4 --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_s.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"render_kernel|primary_stage" -s 2 -c 2 -o gpurun_out/prof_r1s_c4 $CMD > gpurun_out/ncu_full_s.log 2>&1
tail -2 gpurun_out/plain_s.log | cut -c1-400
cat gpurun_out/summary_s.txt
