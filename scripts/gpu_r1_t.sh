#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_t.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary_t.txt
tail -3 gpurun_out/pytest_gpu_t.log
for cfg in C4 C2; do
  extra=""; [ $cfg = C4 ] && extra="--spp 16"
  timeout 300 python bench.py --config $cfg $extra --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$cfg', round(d['value'],1),'Msamples/s', round(d['mrays_s'],1),'Mrays/s box/ray', round(r['box_tests_per_ray'],2), 'sph/ray', round(r['sphere_tests_per_ray'],2))" >> gpurun_out/variants_t.txt
done
cat gpurun_out/variants_t.txt; cat gpurun_out/summary_t.txt
