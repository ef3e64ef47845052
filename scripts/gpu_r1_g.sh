#!/bin/bash
# 2-GPU pass: sample-split with one NCCL reduce, launched the way the driver launches it.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; echo "n2 rc=$?" > gpurun_out/summary_g.txt
tail -5 gpurun_out/bench_c2_n2.err
cat gpurun_out/bench_c2_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?" >> gpurun_out/summary_g.txt
cat gpurun_out/bench_ref_n2.json
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_n1_g.json 2>/dev/null
python - <<'PY'
import json
a=json.loads(open('gpurun_out/bench_c2_n1_g.json').read().strip().splitlines()[-1]); b=json.loads(open('gpurun_out/bench_c2_n2.json').read().strip().splitlines()[-1])
print('N=1', round(a['value'],1), 'N=2', round(b['value'],1), 'efficiency', round(b['value']/(2*a['value']),3), 'e2e', round(a['e2e']['value'],1), round(b['e2e']['value'],1))
PY
cat gpurun_out/summary_g.txt
