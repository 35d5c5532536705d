#!/bin/bash
# round-2 GPU pass B: full parity suite (no -x), halo-stage / register-cap A/B
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2b_pytest.log | tail -30
for cfg in "default" "ast2" "lb1"; do
  case $cfg in
    default) envs="";;
    ast2) envs="SUNET_HALO2_AST=2";;
    lb1) envs="SUNET_LIB=$PWD/exp/libsunet_lb1.so";;
  esac
  echo "=== $cfg ($envs)" | tee -a gpurun_out/r2b_timing.log
  env $envs python scripts/bnb_timing.py 128 >> gpurun_out/r2b_timing.log 2>&1
  env $envs python bench.py --steps 10 --warmup 3 --no-stock --no-eval --no-cpu-baseline --no-u8 > gpurun_out/r2b_bench_$cfg.json 2> gpurun_out/r2b_bench_$cfg.err
  python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_$cfg.json')); print('$cfg', round(d['ms_per_step'],3),'ms/step', round(d['value'],1),'patches/s', d['clocks'], 'G1', round(d['roofline']['achieved'],1), 'G2', round(d['roofline']['wgrad_gemm']['achieved'],1))" | tee -a gpurun_out/r2b_timing.log
done
cat gpurun_out/r2b_timing.log
echo "=== overlap probe (default lib: halo2 capped at 168 registers)" | tee -a gpurun_out/r2b_timing.log
python scripts/overlap_probe.py 64 2>&1 | tee -a gpurun_out/r2b_overlap.log
echo "=== overlap probe (lb1: uncapped registers)" | tee -a gpurun_out/r2b_overlap.log
SUNET_LIB=$PWD/exp/libsunet_lb1.so python scripts/overlap_probe.py 64 2>&1 | tee -a gpurun_out/r2b_overlap.log
