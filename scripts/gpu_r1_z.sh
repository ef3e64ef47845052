#!/bin/bash
# Centre/half-extent slab test vs (min, max) slab test: parity tests, then A/B on the same box.
set -u
mkdir -p gpurun_out; rm -f gpurun_out/variants_z.txt
L=raytracer_go_b200/csrc
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_z.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_z.txt
cp $L/librt_b200.so /tmp/new.so
run() { label="$1"; shift
  env "$@" timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', d['config']['workload'][:3], round(d['value'],1),'Msamples/s ms/step', round(d['ms_per_step'],2), 'box/ray', round(r['box_tests_per_ray'],2))" >> gpurun_out/variants_z.txt 2>&1
}
for rep in 1 2; do
  cp /tmp/new.so $L/librt_b200.so;      run "centre-half C2" RT_B200_PASS_BALANCE=0
  cp $L/ab_minmax.so $L/librt_b200.so;  run "min-max     C2" RT_B200_PASS_BALANCE=0
done
runc() { label="$1"; cfg="$2"
  RT_B200_PASS_BALANCE=0 timeout 300 python bench.py --config $cfg --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', d['config']['workload'][:3], round(d['value'],1),'Msamples/s ms/step', round(d['ms_per_step'],2), 'box/ray', round(r['box_tests_per_ray'],2))" >> gpurun_out/variants_z.txt 2>&1
}
for c in C3 C4 CB; do
  cp /tmp/new.so $L/librt_b200.so;      runc "centre-half" $c
  cp $L/ab_minmax.so $L/librt_b200.so;  runc "min-max    " $c
done
cp /tmp/new.so $L/librt_b200.so
cat gpurun_out/variants_z.txt
