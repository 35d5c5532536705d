#!/bin/bash
# round-2 prologue fusion A/B: full step with SUNET_FUSE_PROLOGUE=0/1 (alternating, same box) + per-kernel times
set -u
mkdir -p gpurun_out
for f in 0 1 0 1 0 1; do
  SUNET_FUSE_PROLOGUE=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-stock --no-eval --no-dp-parity \
    > gpurun_out/r2p_bench_pro$f.json 2> gpurun_out/r2p_bench_pro$f.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2p_bench_pro$f.json").read().strip().splitlines()[-1])
    print("pro=$f", round(d["ms_per_step"], 3), round(d["value"], 1), d["clocks"]["sm_mhz"], d["roofline"].get("tensor"))
except Exception as e:
    print("pro=$f no json", e)
PY
done
