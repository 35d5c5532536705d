#!/bin/bash
# round-2 pass V: re-freeze after the pooled BatchNorm kernel changes — parity suite, bench line, ncu launch lists
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2v_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2v_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2v_pytest.log | tail -20
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1]); print('bench', round(d['ms_per_step'],3),'ms/step', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'], 'eval', round(d['eval']['value'],1), 'G1', round(d['roofline']['achieved'],1), 'G2', round(d['roofline']['wgrad_gemm']['achieved'],1), 'traffic', d['roofline']['traffic'])"
B="--no-stock --no-eval --no-cpu-baseline --no-u8"
python bench.py --steps 20 --warmup 5 --batch 16 $B > gpurun_out/r2v_bench16.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r2v_bench16.json').read().strip().splitlines()[-1]); print('batch16', round(d['ms_per_step'],3),'ms/step', d['clocks'])"
python bench.py --steps 2 --warmup 3 $B > gpurun_out/r2v_plain128.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2v_b128.csv python bench.py --steps 2 --warmup 3 $B > gpurun_out/r2v_ncu128.log 2>&1
echo "ncu128 rc=$?"
python bench.py --steps 2 --warmup 3 --batch 16 $B > gpurun_out/r2v_plain16.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2v_b16.csv python bench.py --steps 2 --warmup 3 --batch 16 $B > gpurun_out/r2v_ncu16.log 2>&1
echo "ncu16 rc=$?"
python scripts/ew_bw.py 128 > gpurun_out/r2v_ew_bw.log 2>&1; echo "ew_bw rc=$?"
