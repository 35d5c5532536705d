#!/bin/bash
# round-2 GPU pass H: frozen kernels — full parity suite, heads bandwidth, default bench line, ncu launch list + heads ncu
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2h_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2h_pytest.log | tail -20
python scripts/ew_bw.py 128 > gpurun_out/r2h_ew_bw.log 2>&1; grep -i "heads" gpurun_out/r2h_ew_bw.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json')); print('bench', round(d['ms_per_step'],3),'ms/step', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'], 'launches/step', d['gpu_launches']/d['steps'], 'eval', round(d['eval']['value'],1), 'G1', round(d['roofline']['achieved'],1), 'G2', round(d['roofline']['wgrad_gemm']['achieved'],1))"
python __graft_entry__.py smoke 2>&1 | tail -2
B="--no-stock --no-eval --no-cpu-baseline --no-u8"
python bench.py --steps 2 --warmup 3 $B > gpurun_out/r2h_plain128.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2h_b128.csv python bench.py --steps 2 --warmup 3 $B > gpurun_out/r2h_ncu128.log 2>&1
echo "ncu128 rc=$?"
python scripts/ncu_target.py headsbwd > gpurun_out/plain_headsbwd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:heads_bwd_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02h_headsbwd \
    python scripts/ncu_target.py headsbwd > gpurun_out/ncu_headsbwd.log 2>&1
echo "headsbwd rc=$?"
