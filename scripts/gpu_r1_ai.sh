#!/bin/bash
# Parallel host BVH build on the GPU box: tests, create/render/destroy breakdown (C2, C3, C4; C4 also
# single-threaded for the before/after), bench C2 and C4.
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_ai.txt
echo "nproc $(nproc)" | tee $S
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ai.txt 2>&1; echo "pytest rc=$?" | tee -a $S
tail -2 gpurun_out/pytest_ai.txt | tee -a $S
for cfg in C2 C3 C4; do timeout 300 python scripts/e2e_breakdown.py $cfg 2>&1 | tee -a $S; done
echo "--- C4, RT_B200_BVH_THREADS=1" | tee -a $S
RT_B200_BVH_THREADS=1 RT_B200_BVH_TIMING=1 timeout 300 python scripts/e2e_breakdown.py C4 2>&1 | tail -12 | tee -a $S
echo "--- C4, default threads, phases" | tee -a $S
RT_B200_BVH_TIMING=1 timeout 300 python scripts/e2e_breakdown.py C4 2>&1 | tail -12 | tee -a $S
timeout 600 python bench.py > gpurun_out/bench_ai_C2.json 2> gpurun_out/bench_ai.err; echo "bench rc=$?" | tee -a $S
timeout 600 python bench.py --config C4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ai_C4.json 2>> gpurun_out/bench_ai.err; echo "C4 rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import json
for c in ['C2','C4']:
    try:
        d=json.loads(open(f'gpurun_out/bench_ai_{c}.json').read().strip().splitlines()[-1]); r=d['roofline']
        print(c, d['config']['workload'], '|', round(d['value'],1),'Msamples/s', round(d['mrays_s'],1),'Mrays/s e2e', round(d['e2e']['value'],1), 'frac', round(r['frac'],4))
    except Exception as e: print(c, 'ERR', e)
PY
