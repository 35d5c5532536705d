#!/bin/bash
# round-2 GPU pass K: frozen kernels — parity suite, bench lines (batch 128 / 16), ncu launch lists, ncu --set full of the GEMMs
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2k_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2k_pytest.log | tail -20
python scripts/ew_bw.py 128 > gpurun_out/r2k_ew_bw.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2k_bench.json')); print('bench', round(d['ms_per_step'],3),'ms/step', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'], 'eval', round(d['eval']['value'],1), 'G1', round(d['roofline']['achieved'],1), 'G2', round(d['roofline']['wgrad_gemm']['achieved'],1))"
B="--no-stock --no-eval --no-cpu-baseline --no-u8"
python bench.py --steps 20 --warmup 5 --batch 16 $B > gpurun_out/r2k_bench16.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2k_bench16.json')); print('batch16', round(d['ms_per_step'],3),'ms/step', d['clocks'])"
python bench.py --steps 2 --warmup 3 $B > gpurun_out/r2k_plain128.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2k_b128.csv python bench.py --steps 2 --warmup 3 $B > gpurun_out/r2k_ncu128.log 2>&1
echo "ncu128 rc=$?"
python bench.py --steps 2 --warmup 3 --batch 16 $B > gpurun_out/r2k_plain16.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2k_b16.csv python bench.py --steps 2 --warmup 3 --batch 16 $B > gpurun_out/r2k_ncu16.log 2>&1
echo "ncu16 rc=$?"
bash scripts/ncu_full.sh r02k
run() {
  python scripts/ncu_target.py $1 > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o gpurun_out/prof_r02k_$1 \
      python scripts/ncu_target.py $1 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
run headsbwd heads_bwd_kernel
