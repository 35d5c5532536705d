#!/bin/bash
# ncu --set full on one launch of each dominant kernel (B200_PROFILING.md recipe): plain run first, then profile.
# usage: scripts/ncu_full.sh [tag]   -> gpurun_out/prof_<tag>_<mode>.ncu-rep ; summarise with scripts/ncu_summary.py
set -u
tag=${1:-r01}
mkdir -p gpurun_out
run() {  # mode kernel-regex
  python scripts/ncu_target.py $1 > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o gpurun_out/prof_${tag}_$1 \
      python scripts/ncu_target.py $1 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
run conv256 conv3_halo2_kernel
run conv128 conv3_halo2_kernel
run conv64 conv3_halo2_kernel
run wgrad256 wgrad_gemm
run wgrad128 wgrad_gemm
run wgrad64 wgrad64
run conv256pro conv3_halo2_kernel
run wgrad256pro wgrad_gemm
run bnbwd bn_bwd_apply_flat_kernel
