#!/bin/bash
# round-2 prologue fusion bring-up: kernel-level bit-identity, plan-level bit-identity, then A/B of the full step
set -u
mkdir -p gpurun_out
for g in g1_prologue g2_wgrad_prologue; do
  timeout 180 python scripts/gpu_probe.py $g > gpurun_out/r2o_$g.log 2>&1
  echo "$g rc=$?"; grep -E "BAD|skip|PASS|FAIL|timed out|error" gpurun_out/r2o_$g.log | head -20
done
timeout 600 python -m pytest tests/test_gpu_prologue.py -x -q > gpurun_out/r2o_plan.log 2>&1
echo "plan test rc=$?"; tail -15 gpurun_out/r2o_plan.log
for f in 0 1 0 1; do
  SUNET_FUSE_PROLOGUE=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-stock --no-eval --no-dp-parity \
    > gpurun_out/r2o_bench_pro$f.json 2> gpurun_out/r2o_bench_pro$f.err
  echo "pro=$f rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2o_bench_pro$f.json").read().strip().splitlines()[-1])
    print("pro=$f", d["ms_per_step"], d["value"], d["clocks"])
except Exception as e:
    print("no json", e)
PY
done
