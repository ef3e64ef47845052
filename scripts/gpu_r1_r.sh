#!/bin/bash
# Two up-front rejection trials: parity + speed.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cp raytracer_go_b200/csrc/librt_b200.so gpurun_out/librt_b200_r1r.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary_r.txt
tail -3 gpurun_out/pytest_gpu_r.log
for i in 1 2; do
timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2', round(d['value'],1),'Msamples/s', round(d['mrays_s'],1),'Mrays/s ms/step', round(d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],4))" >> gpurun_out/variants_r.txt
done
timeout 200 python bench.py --config CB --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cornell', round(d['value'],1), round(d['mrays_s'],1))" >> gpurun_out/variants_r.txt
cat gpurun_out/variants_r.txt; cat gpurun_out/summary_r.txt
