#!/bin/bash
# round-2: full GPU suite with prologue fusion on by default, then the N=1 bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2r_pytest.log
timeout 900 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r2r_bench.json
