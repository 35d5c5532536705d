#!/bin/bash
# round-2 GPU pass D: parity suite, stream-kernel bandwidths, default bench line, configs[4] sweep
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2d_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2d_pytest.log | tail -30
python scripts/ew_bw.py 128 > gpurun_out/r2d_ew_bw.log 2>&1; tail -12 gpurun_out/r2d_ew_bw.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2d_bench.json')); print('bench', round(d['ms_per_step'],3),'ms/step', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'], 'launches/step', d['gpu_launches']/d['steps'], 'eval', round(d['eval']['value'],1), 'hist GB/s', round(d['eval']['metric_hist']['achieved'],1))"
timeout 2400 bash scripts/batch_sweep.sh gpurun_out/r2d_batch_sweep.txt
