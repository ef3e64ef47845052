#!/bin/bash
# First GPU pass: smoke, gpu tests, bench, variants, ncu launch list + full capture.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench_c2.json
# variants (device-only, short)
for blk in 256 512 1024; do for leaf in 1 2 4; do
  RT_B200_BLOCK=$blk RT_B200_MAX_LEAF=$leaf timeout 300 python bench.py --spp 100 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('blk $blk leaf $leaf', round(d['value'],1),'Msamples/s', round(d['mrays_s'],1),'Mrays/s box/ray',round(d['roofline']['box_tests_per_ray'],2),'sph/ray',round(d['roofline']['sphere_tests_per_ray'],2))" >> gpurun_out/variants.txt 2>&1
done; done
RT_B200_NO_SMEM=1 timeout 300 python bench.py --spp 100 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nosmem blk256 leaf4', round(d['value'],1),'Msamples/s', round(d['mrays_s'],1))" >> gpurun_out/variants.txt 2>&1
cat gpurun_out/variants.txt
# ncu: launch list, then one full capture of the megakernel (same command line, plain run first)
CMD="python bench.py --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1a.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/prof_r1a $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
