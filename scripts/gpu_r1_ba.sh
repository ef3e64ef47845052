#!/bin/bash
# Validation of the final round-1 configuration (512 x 2 CTAs, threaded BVH build): tests in both kernel modes,
# smoke, bench on every config, reference arm, launch list of the bench command.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ba.txt 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary_ba.txt
RT_B200_KERNEL=mega timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ba_mega.txt 2>&1; echo "pytest mega rc=$?" | tee -a gpurun_out/summary_ba.txt
python __graft_entry__.py smoke > gpurun_out/smoke_ba.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary_ba.txt; tail -1 gpurun_out/smoke_ba.log
timeout 600 python bench.py > gpurun_out/bench_ba_C2.json 2> gpurun_out/bench_ba.err; echo "bench rc=$?" | tee -a gpurun_out/summary_ba.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ba_reference.json 2>/dev/null; echo "ref rc=$?" | tee -a gpurun_out/summary_ba.txt
for cfg in C1 C3 C4 C5 CB; do
  extra=""; [ $cfg = C5 ] && extra="--spp 64"
  timeout 600 python bench.py --config $cfg $extra --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ba_${cfg}.json 2>> gpurun_out/bench_ba.err; echo "$cfg rc=$?" | tee -a gpurun_out/summary_ba.txt
done
python - <<'PY' | tee -a gpurun_out/summary_ba.txt
import json
for c in ['C2','C1','C3','C4','C5','CB']:
    try:
        d=json.loads(open(f'gpurun_out/bench_ba_{c}.json').read().strip().splitlines()[-1]); r=d['roofline']
        print(c, d['config']['workload'], '|', round(d['value'],1),'Msamples/s', round(d['mrays_s'],1),'Mrays/s e2e', round(d['e2e']['value'],1), 'frac', round(r['frac'],4), 'seg', round(r['segments_per_sample'],2), 'box/ray', round(r['box_tests_per_ray'],1))
    except Exception as e: print(c, 'ERR', e)
PY
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_ba.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_r1ba.csv $CMD > gpurun_out/ncu_launch_ba.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/summary_ba.txt
