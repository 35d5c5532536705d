#!/bin/bash
# one vs two waves of split-K CTAs per weight-gradient launch: step A/B at batch 128 and 16
set -u
mkdir -p gpurun_out
B="--no-stock --no-eval --no-cpu-baseline --no-u8 --no-dp-parity --steps 20 --warmup 5"
for rep in 1 2; do
for cfg in "128 2" "128 1" "16 2" "16 1"; do
  set -- $cfg
  SUNET_WGRAD_WAVES=$2 timeout 600 python bench.py $B --batch $1 > gpurun_out/r2w_b$1_w$2.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2w_b$1_w$2.json").read().strip().splitlines()[-1])
print("batch $1 waves $2:", round(d["ms_per_step"], 3), "ms", d["clocks"]["sm_mhz"])
PY
done
done
