#!/bin/bash
# Confirm the new defaults (64 SAH bins, 64 Mi paths per pass): gpu tests, bench on every config.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v.txt 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_v.txt
tail -3 gpurun_out/pytest_v.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo "bench rc=$?"
cat gpurun_out/bench_v.json
for c in C1 C3 C4 CB; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v_$c.json 2>> gpurun_out/bench_v.err; echo "$c rc=$?"
  cat gpurun_out/bench_v_$c.json
done
