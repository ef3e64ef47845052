#!/bin/bash
# gpu_run.sh — the one GPU-side runner (replaces round 1's 50 one-off scripts/gpu_r1_*.sh).
# Called on the B200 box through gpurun, e.g.
#   gpurun --timeout 900 -- 'bash scripts/gpu_run.sh TAG validate bench ncu-launches'
# Every step writes to gpurun_out/<TAG>_*; steps run in the order given.
#
#   validate        pytest -m gpu (default two-stage mode) + smoke()
#   validate-mega   pytest -m gpu with RT_B200_KERNEL=mega (one-stage mode)
#   validate-debug  pytest -m gpu with RT_B200_DEBUG=1 (bounds-checked kernel instantiations)
#   bench           bench.py with the driver's default flags
#   bench-short     bench.py --steps 5 --warmup 3 --no-cpu-baseline
#   reference       bench.py --impl reference --steps 2 --warmup 1
#   configs         bench.py on C1 C3 C4 C5(64 spp) CB, short
#   sweep:VAR:a,b,c bench-short once per value of environment variable VAR
#   ab:NAME[:cfgs]  the default library against the variant build csrc/librt_b200_NAME.so
#   ncu-launches    per-launch durations of the bench command (gpu__time_duration.sum)
#   ncu-full[:SPP[:CFG]]  ncu --set full of one primary + one secondary launch at SPP (default 82) spp of config CFG (default C2)
#   scale           bench.py at N = 1/2/4/8 under torchrun as the driver launches it, weak + strong
#   multi           rt_render_multi on C2 / C5 in both split modes on all GPUs of the box
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG="$1"; shift
S="gpurun_out/${TAG}_summary.txt"; : > "$S"
say() { echo "$@" | tee -a "$S"; }
line() { python - "$@" <<'PY' | tee -a "$S"
import json, sys
label, path = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1]); r = d.get('roofline', {})
    e = d.get('e2e') or {}
    print(label, d['config']['workload'], '|', round(d['value'], 1), d['unit'], round(d.get('mrays_s', 0), 1), 'Mrays/s e2e',
          round(e.get('value', 0), 1), 'ms/step', round(d['ms_per_step'], 2), 'frac', round(r.get('frac', 0), 4),
          'box/ray', round(r.get('box_tests_per_ray', 0), 2), 'sph/ray', round(r.get('sphere_tests_per_ray', 0), 2),
          'clk', d.get('clocks', {}).get('sm_mhz'))
except Exception as ex:
    print(label, 'ERR', ex)
PY
}
NG=$(nvidia-smi -L | wc -l)
say "tag=$TAG gpus=$NG $(nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -1)"
for step in "$@"; do
  case "$step" in
    validate)
      timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; say "pytest rc=$? $(tail -1 gpurun_out/${TAG}_pytest.txt)"
      python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; say "smoke rc=$? $(tail -1 gpurun_out/${TAG}_smoke.log)" ;;
    validate-mega)
      RT_B200_KERNEL=mega timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_mega.txt 2>&1; say "pytest mega rc=$? $(tail -1 gpurun_out/${TAG}_pytest_mega.txt)" ;;
    validate-debug)
      RT_B200_DEBUG=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_debug.txt 2>&1; say "pytest debug rc=$? $(tail -1 gpurun_out/${TAG}_pytest_debug.txt)" ;;
    bench)
      timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; say "bench rc=$?"; line C2 gpurun_out/${TAG}_bench.json ;;
    bench-short)
      timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_short.json 2> gpurun_out/${TAG}_bench_short.err; say "bench-short rc=$?"; line C2 gpurun_out/${TAG}_bench_short.json ;;
    reference)
      timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_reference.json 2>/dev/null; say "reference rc=$?"; line REF gpurun_out/${TAG}_reference.json ;;
    configs)
      for cfg in C1 C3 C4 C5 CB; do
        extra=""; [ $cfg = C5 ] && extra="--spp 64"
        timeout 600 python bench.py --config $cfg $extra --steps 3 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/${TAG}_bench_${cfg}.json 2>> gpurun_out/${TAG}_configs.err
        line $cfg gpurun_out/${TAG}_bench_${cfg}.json
      done ;;
    sweep:*)
      var=$(echo "$step" | cut -d: -f2); vals=$(echo "$step" | cut -d: -f3 | tr ',' ' ')
      cfgs=$(echo "$step" | cut -d: -f4 | tr ',' ' '); [ -z "$cfgs" ] && cfgs=C2
      for v in $vals; do for cfg in $cfgs; do
        env $var=$v timeout 600 python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-strong > gpurun_out/${TAG}_sweep_${var}_${v}_${cfg}.json 2>> gpurun_out/${TAG}_sweep.err
        line "$var=$v" gpurun_out/${TAG}_sweep_${var}_${v}_${cfg}.json
      done; done ;;
    ab:*)  # ab:NAME[:cfgs] — the default library against csrc/librt_b200_NAME.so (csrc/Makefile `variant`)
      name=$(echo "$step" | cut -d: -f2); cfgs=$(echo "$step" | cut -s -d: -f3 | tr ',' ' '); [ -z "$cfgs" ] && cfgs="C2"
      for cfg in $cfgs; do for which in default $name; do
        libenv=""; [ $which != default ] && libenv="RT_B200_LIBRARY=raytracer_go_b200/csrc/librt_b200_${which}.so"
        env $libenv timeout 600 python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-strong > gpurun_out/${TAG}_ab_${which}_${cfg}.json 2>> gpurun_out/${TAG}_ab.err
        line "lib=$which" gpurun_out/${TAG}_ab_${which}_${cfg}.json
      done; done ;;
    ncu-launches)
      CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-strong"
      $CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
      say "ncu-launches rc=$?" ;;
    ncu-full*)
      spp=$(echo "$step" | cut -s -d: -f2); [ -z "$spp" ] && spp=82
      cfg=$(echo "$step" | cut -s -d: -f3); [ -z "$cfg" ] && cfg=C2
      CMD="python bench.py --config $cfg --spp $spp --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-strong"
      $CMD > gpurun_out/${TAG}_plain_full.log 2>&1 && \
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"render_kernel|primary_stage" -s 2 -c 2 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
      say "ncu-full rc=$?" ;;
    scale)
      for split in weak strong; do for n in 1 2 4 8; do
        [ $n -gt $NG ] && break
        out=gpurun_out/${TAG}_scale_${split}_n$n.json
        if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --split $split --steps 5 --warmup 3 --no-cpu-baseline > $out 2> gpurun_out/${TAG}_scale.err
        else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --split $split --steps 5 --warmup 3 --no-cpu-baseline > $out 2>> gpurun_out/${TAG}_scale.err; fi
        line "$split N=$n" $out
      done; done ;;
    multi)
      timeout 900 python scripts/render_multi_timing.py 2>&1 | tee -a "$S" ;;
    *) say "unknown step $step" ;;
  esac
done
