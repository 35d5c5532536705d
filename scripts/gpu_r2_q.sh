#!/bin/bash
# same-box A/B: the previous commit (copy under _ab/) against the working tree, SUNET_FUSE_PROLOGUE=0
set -u
mkdir -p gpurun_out
ROOT=$(pwd)
for t in head tree head tree; do
  if [ $t = head ]; then cd $ROOT/_ab; else cd $ROOT; fi
  SUNET_FUSE_PROLOGUE=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-stock --no-eval --no-dp-parity \
    > $ROOT/gpurun_out/r2q_$t.json 2> $ROOT/gpurun_out/r2q_$t.err
  cd $ROOT
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2q_$t.json").read().strip().splitlines()[-1])
    print("$t", round(d["ms_per_step"], 3), round(d["value"], 1), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$t no json", e)
PY
done
