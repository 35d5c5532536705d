#!/bin/bash
# round-2 GPU pass C: full parity suite incl. fp32 check mode; ncu launch lists at batch 128 and batch 16
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2c_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2c_pytest.log | tail -30
B="--no-stock --no-eval --no-cpu-baseline --no-u8"
python bench.py --steps 2 --warmup 3 $B > gpurun_out/r2c_plain128.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2c_b128.csv python bench.py --steps 2 --warmup 3 $B > gpurun_out/r2c_ncu128.log 2>&1
echo "ncu128 rc=$?"
python bench.py --steps 2 --warmup 3 --batch 16 $B > gpurun_out/r2c_plain16.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2c_b16.csv python bench.py --steps 2 --warmup 3 --batch 16 $B > gpurun_out/r2c_ncu16.log 2>&1
echo "ncu16 rc=$?"
python bench.py --steps 20 --warmup 5 --batch 16 $B > gpurun_out/r2c_bench16.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2c_bench16.json')); print('batch16', round(d['ms_per_step'],3),'ms/step', d['clocks'], d['gpu_launches']/d['steps'])"
