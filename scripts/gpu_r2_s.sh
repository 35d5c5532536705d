#!/bin/bash
# round-2 GPU pass S: frozen kernels (prologue fusion + wide Cout=64 wgrad) — parity suite, bench lines, ncu launch lists,
# ncu --set full of the GEMM kernels
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2s_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2s_pytest.log | tail -20
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1]); print('bench', round(d['ms_per_step'],3),'ms/step', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'], 'eval', round(d['eval']['value'],1), 'G1', round(d['roofline']['achieved'],1), 'G2', round(d['roofline']['wgrad_gemm']['achieved'],1), 'traffic', d['roofline']['traffic'])"
B="--no-stock --no-eval --no-cpu-baseline --no-u8"
python bench.py --steps 20 --warmup 5 --batch 16 $B > gpurun_out/r2s_bench16.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r2s_bench16.json').read().strip().splitlines()[-1]); print('batch16', round(d['ms_per_step'],3),'ms/step', d['clocks'])"
python bench.py --steps 2 --warmup 3 $B > gpurun_out/r2s_plain128.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2s_b128.csv python bench.py --steps 2 --warmup 3 $B > gpurun_out/r2s_ncu128.log 2>&1
echo "ncu128 rc=$?"
python bench.py --steps 2 --warmup 3 --batch 16 $B > gpurun_out/r2s_plain16.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2s_b16.csv python bench.py --steps 2 --warmup 3 --batch 16 $B > gpurun_out/r2s_ncu16.log 2>&1
echo "ncu16 rc=$?"
bash scripts/ncu_full.sh r02s
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
