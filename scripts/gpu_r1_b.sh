#!/bin/bash
# Second GPU pass: kernel v1 (convergent RNG blocks, while-while traversal, regen threshold).
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/variants_b.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary_b.txt
tail -5 gpurun_out/pytest_gpu_b.log
run() { # label env...
  label="$1"; shift
  env "$@" timeout 300 python bench.py --spp 100 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label', round(d['value'],1),'Msamples/s', round(d['mrays_s'],1),'Mrays/s', 'frac', round(d['roofline']['frac'],4))" >> gpurun_out/variants_b.txt 2>&1
}
run "b256m3 regen1" RT_B200_BLOCK=256 RT_B200_MINB=3
run "b256m3 regen1 again" RT_B200_BLOCK=256 RT_B200_MINB=3
run "b256m2" RT_B200_BLOCK=256 RT_B200_MINB=2
run "b256m4" RT_B200_BLOCK=256 RT_B200_MINB=4
run "b512m1" RT_B200_BLOCK=512 RT_B200_MINB=1
run "b512m2" RT_B200_BLOCK=512 RT_B200_MINB=2
run "b1024m1" RT_B200_BLOCK=1024 RT_B200_MINB=1
for rg in 2 4 8 12 16 24; do run "b256m3 regen$rg" RT_B200_BLOCK=256 RT_B200_MINB=3 RT_B200_REGEN_MIN=$rg; done
run "b256m3 nosmem" RT_B200_NO_SMEM=1
run "b256m3 leaf1" RT_B200_MAX_LEAF=1
run "b256m3 leaf2" RT_B200_MAX_LEAF=2
run "b256m3 leaf8" RT_B200_MAX_LEAF=8
cat gpurun_out/variants_b.txt
python - <<'PY' > gpurun_out/e2e_breakdown.txt 2>&1
import time, numpy as np
from raytracer_go_b200 import api, scenes
s = scenes.random_scene(); cam = api.camera_from_options(scenes.camera_options(1200, 500))
for it in range(3):
    t0=time.perf_counter(); sc = api.Scene(s); t1=time.perf_counter()
    rgb,_,st = sc.render(cam); t2=time.perf_counter()
    sc.close(); t3=time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms render {1e3*(t2-t1):.1f} ms (device {st.ms_render:.1f}, megakernel {st.ms_megakernel:.1f}, lib total {st.ms_total:.1f}) destroy {1e3*(t3-t2):.1f} ms")
PY
cat gpurun_out/e2e_breakdown.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c2_b.json 2> gpurun_out/bench_c2_b.err; echo "bench rc=$?" >> gpurun_out/summary_b.txt
cat gpurun_out/bench_c2_b.json
CMD="python bench.py --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/prof_r1b $CMD > gpurun_out/ncu_full_b.log 2>&1
ls -la gpurun_out
