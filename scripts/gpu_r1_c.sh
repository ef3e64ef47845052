#!/bin/bash
# Third GPU pass: workspace cache, wide-multiply Philox, FMNMX3, regen_min 8.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cp raytracer_go_b200/csrc/librt_b200.so gpurun_out/librt_b200_r1c.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary_c.txt
tail -5 gpurun_out/pytest_gpu_c.log
python - <<'PY' > gpurun_out/e2e_breakdown_c.txt 2>&1
import time, numpy as np
from raytracer_go_b200 import api, scenes
s = scenes.random_scene(); cam = api.camera_from_options(scenes.camera_options(1200, 500))
for it in range(5):
    t0=time.perf_counter(); sc = api.Scene(s); t1=time.perf_counter()
    rgb,_,st = sc.render(cam); t2=time.perf_counter()
    sc.close(); t3=time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms render {1e3*(t2-t1):.1f} ms (device {st.ms_render:.1f}, megakernel {st.ms_megakernel:.1f}, lib total {st.ms_total:.1f}) destroy {1e3*(t3-t2):.1f} ms")
PY
cat gpurun_out/e2e_breakdown_c.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2_c.json 2> gpurun_out/bench_c2_c.err; echo "bench rc=$?" >> gpurun_out/summary_c.txt
cat gpurun_out/bench_c2_c.json
for rg in 1 4 8 16; do
RT_B200_REGEN_MIN=$rg timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('regen $rg', round(d['value'],1),'Msamples/s frac', round(d['roofline']['frac'],4), 'share', round(d['roofline']['kernel_share_of_step'],3))" >> gpurun_out/variants_c.txt
done
cat gpurun_out/variants_c.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_c.json 2>&1
cat gpurun_out/bench_ref_c.json
CMD="python bench.py --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1c.csv $CMD > gpurun_out/ncu_launch_c.log 2>&1
$CMD > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/prof_r1c $CMD > gpurun_out/ncu_full_c.log 2>&1
ls -la gpurun_out | tail -20
