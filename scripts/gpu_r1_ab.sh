#!/bin/bash
# A/B on one box: new library vs the previous commit (ab_prev.so).
set -u
mkdir -p gpurun_out; rm -f gpurun_out/variants_ab.txt
L=raytracer_go_b200/csrc
cp $L/librt_b200.so /tmp/new.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ab.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_ab.txt
runc() { label="$1"; cfg="$2"
  RT_B200_PASS_BALANCE=0 timeout 300 python bench.py --config $cfg --steps 4 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', d['config']['workload'][:3], round(d['value'],1),'Msamples/s ms/step', round(d['ms_per_step'],2), 'frac', round(r['frac'],4))" >> gpurun_out/variants_ab.txt 2>&1
}
for rep in 1 2; do
  for c in C2 CB C1; do
    cp /tmp/new.so $L/librt_b200.so;     runc "new " $c
    cp $L/ab_prev.so $L/librt_b200.so;   runc "prev" $c
  done
done
cp /tmp/new.so $L/librt_b200.so
cat gpurun_out/variants_ab.txt
