#!/bin/bash
# Balanced passes: tests in both kernel modes, bench, launch list of the same command.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_x.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_x.txt
RT_B200_KERNEL=mega timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_x_mega.txt 2>&1; echo "pytest mega rc=$?"; tail -2 gpurun_out/pytest_x_mega.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_x.json
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_x.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_r1x.csv $CMD > gpurun_out/ncu_launch_x.log 2>&1
echo "ncu rc=$?"
