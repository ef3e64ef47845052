#!/bin/bash
# Sixth GPU pass: config-sized tests, reverted traversal.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cp raytracer_go_b200/csrc/librt_b200.so gpurun_out/librt_b200_r1f.so
timeout 1200 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/pytest_gpu_f.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary_f.txt
tail -40 gpurun_out/pytest_gpu_f.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c2_f.json 2>gpurun_out/bench_c2_f.err; echo "bench rc=$?" >> gpurun_out/summary_f.txt
python -c "
import json; d=json.loads(open('gpurun_out/bench_c2_f.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['frac'])"
cat gpurun_out/summary_f.txt
