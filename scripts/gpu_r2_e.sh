#!/bin/bash
# round-2 GPU pass E: ncu --set full on the heads / loss / metric / Adam kernels, stream-kernel bandwidth log
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -q -k "loss or heads or trainer" > gpurun_out/r2e_pytest.log 2>&1; tail -3 gpurun_out/r2e_pytest.log
python scripts/ew_bw.py 128 > gpurun_out/r2e_ew_bw.log 2>&1; tail -8 gpurun_out/r2e_ew_bw.log
run() {  # mode kernel-regex
  python scripts/ncu_target.py $1 > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o gpurun_out/prof_r02_$1 \
      python scripts/ncu_target.py $1 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
run headsbwd heads_bwd_kernel
run headsfwd bn_relu_heads_kernel
run hist metric_hist_kernel
run losssums loss_sums_kernel
run lossbwd loss_bwd_kernel
run adam adam_kernel
