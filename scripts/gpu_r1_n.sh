#!/bin/bash
# Final validation of the default (two-stage) configuration: tests, smoke, bench, reference arm, ncu evidence.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cp raytracer_go_b200/csrc/librt_b200.so gpurun_out/librt_b200_r1n.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_n.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary_n.txt
RT_B200_KERNEL=mega timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_n_mega.log 2>&1; echo "pytest mega rc=$?" >> gpurun_out/summary_n.txt
tail -3 gpurun_out/pytest_gpu_n.log
python __graft_entry__.py smoke > gpurun_out/smoke_n.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary_n.txt; tail -1 gpurun_out/smoke_n.log
timeout 600 python bench.py > gpurun_out/bench_c2_n.json 2> gpurun_out/bench_c2_n.err; echo "bench rc=$?" >> gpurun_out/summary_n.txt
cat gpurun_out/bench_c2_n.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n.json 2>/dev/null; echo "ref rc=$?" >> gpurun_out/summary_n.txt
for cfg in C1 C3 C4 C5 CB; do
  extra=""; [ $cfg = C5 ] && extra="--spp 64"; [ $cfg = C4 ] && extra="--spp 16"
  timeout 600 python bench.py --config $cfg $extra --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${cfg}_n.json 2> gpurun_out/bench_${cfg}_n.err; echo "$cfg rc=$?" >> gpurun_out/summary_n.txt
done
python - <<'PY'
import json
for c in ['c2','C1','C3','C4','C5','CB']:
    try:
        d=json.loads(open(f'gpurun_out/bench_{c}_n.json').read().strip().splitlines()[-1]); r=d['roofline']
        print(c, d['config']['workload'], '|', round(d['value'],1),'Msamples/s', round(d['mrays_s'],1),'Mrays/s e2e', round(d['e2e']['value'],1), 'frac', round(r['frac'],4), 'seg', round(r['segments_per_sample'],2), 'box/ray', round(r['box_tests_per_ray'],1))
    except Exception as e: print(c, 'ERR', e)
PY
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_n.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_r1n.csv $CMD > gpurun_out/ncu_launch_n.log 2>&1
CMD2="python bench.py --spp 39 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/plain_n2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"render_kernel|primary_stage" -s 2 -c 2 -o gpurun_out/prof_r1n $CMD2 > gpurun_out/ncu_full_n.log 2>&1
cat gpurun_out/summary_n.txt
