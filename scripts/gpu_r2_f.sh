#!/bin/bash
# round-2 GPU pass F: A-collector reuse in the weight-gradient GEMMs (A/B + parity + step time)
set -u
mkdir -p gpurun_out
python scripts/wgrad_reuse_ab.py > gpurun_out/r2f_wgrad_reuse.log 2>&1; cat gpurun_out/r2f_wgrad_reuse.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize_kernels.py tests/test_gpu_fullsize.py -m gpu -q -k "wgrad or fused_and_unfused" > gpurun_out/r2f_pytest.log 2>&1; tail -4 gpurun_out/r2f_pytest.log
B="--no-stock --no-eval --no-cpu-baseline --no-u8 --steps 10 --warmup 3"
for r in 0 1; do
  SUNET_WGRAD_A_REUSE=$r python bench.py $B > gpurun_out/r2f_bench_reuse$r.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_reuse$r.json')); print('reuse=$r', round(d['ms_per_step'],3),'ms/step', round(d['value'],1),'patches/s', d['clocks'], 'G1', round(d['roofline']['achieved'],1), 'G2', round(d['roofline']['wgrad_gemm']['achieved'],1))" | tee -a gpurun_out/r2f_wgrad_reuse.log
done
