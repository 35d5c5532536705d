#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_unet_ce.py -m gpu -q > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2i_pytest.log
python scripts/ew_bw.py 128 > gpurun_out/r2i_ew_bw.log 2>&1; grep -i "heads" gpurun_out/r2i_ew_bw.log
python scripts/ncu_target.py headsbwd > gpurun_out/plain_headsbwd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:heads_bwd_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02i_headsbwd \
    python scripts/ncu_target.py headsbwd > gpurun_out/ncu_headsbwd.log 2>&1
echo "headsbwd rc=$?"
python bench.py --steps 10 --warmup 3 --no-stock --no-cpu-baseline --no-u8 > gpurun_out/r2i_bench.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2i_bench.json')); print('bench', round(d['ms_per_step'],3),'ms/step', round(d['value'],1), d['clocks'], 'eval', round(d['eval']['value'],1))"
