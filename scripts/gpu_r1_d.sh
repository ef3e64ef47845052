#!/bin/bash
# Fourth GPU pass: pool-variant megakernel (per-warp ray pool, dynamic fetch) vs lock-step megakernel.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cp raytracer_go_b200/csrc/librt_b200.so gpurun_out/librt_b200_r1d.so
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_d_mega.log 2>&1; echo "pytest mega rc=$?" > gpurun_out/summary_d.txt
RT_B200_KERNEL=pool timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_d_pool.log 2>&1; echo "pytest pool rc=$?" >> gpurun_out/summary_d.txt
tail -15 gpurun_out/pytest_gpu_d_pool.log
cat gpurun_out/summary_d.txt
run() { label="$1"; shift
  env "$@" timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label', round(d['value'],1),'Msamples/s', round(d['mrays_s'],1),'Mrays/s frac', round(d['roofline']['frac'],4), 'share', round(d['roofline']['kernel_share_of_step'],3))" >> gpurun_out/variants_d.txt 2>&1
}
run "mega b256m3" RT_B200_KERNEL=mega
run "pool b512 k4" RT_B200_KERNEL=pool RT_B200_POOL_BLOCK=512 RT_B200_POOL_K=4
run "pool b512 k2" RT_B200_KERNEL=pool RT_B200_POOL_BLOCK=512 RT_B200_POOL_K=2
run "pool b512 k8" RT_B200_KERNEL=pool RT_B200_POOL_BLOCK=512 RT_B200_POOL_K=8
run "pool b768 k2" RT_B200_KERNEL=pool RT_B200_POOL_BLOCK=768 RT_B200_POOL_K=2
run "pool b768 k3" RT_B200_KERNEL=pool RT_B200_POOL_BLOCK=768 RT_B200_POOL_K=3
run "pool b1024 k2" RT_B200_KERNEL=pool RT_B200_POOL_BLOCK=1024 RT_B200_POOL_K=2
run "pool b256 k4" RT_B200_KERNEL=pool RT_B200_POOL_BLOCK=256 RT_B200_POOL_K=4
run "pool b512 k4 nosmem" RT_B200_KERNEL=pool RT_B200_POOL_BLOCK=512 RT_B200_POOL_K=4 RT_B200_NO_SMEM=1
cat gpurun_out/variants_d.txt
CMD="python bench.py --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
RT_B200_KERNEL=pool $CMD > gpurun_out/plain_d.log 2>&1 && \
RT_B200_KERNEL=pool ncu --set full --clock-control none --import-source on -k regex:render_pool_kernel -s 1 -c 1 -o gpurun_out/prof_r1d $CMD > gpurun_out/ncu_full_d.log 2>&1
ls -la gpurun_out | tail -8
