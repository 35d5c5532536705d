#!/bin/bash
# round-2 GPU pass J: two epilogue groups in the halo conv kernel (A/B against the 4-warp epilogue in exp/)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r2j_pytest_k.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/r2j_pytest_k.log
for cfg in epi2 epi1; do
  envs=""; [ $cfg = epi1 ] && envs="SUNET_LIB=$PWD/exp/libsunet_epi1.so"
  echo "=== $cfg" | tee -a gpurun_out/r2j_timing.log
  env $envs timeout 300 python scripts/bnb_timing.py 128 >> gpurun_out/r2j_timing.log 2>&1
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-stock --no-eval --no-cpu-baseline --no-u8 > gpurun_out/r2j_bench_$cfg.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_$cfg.json')); print('$cfg', round(d['ms_per_step'],3),'ms/step', round(d['value'],1),'patches/s', d['clocks'], 'G1', round(d['roofline']['achieved'],1), 'G2', round(d['roofline']['wgrad_gemm']['achieved'],1))" | tee -a gpurun_out/r2j_timing.log
done
cat gpurun_out/r2j_timing.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1
echo "all rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2j_pytest.log | tail
python scripts/ew_bw.py 128 2>&1 | grep -i heads
