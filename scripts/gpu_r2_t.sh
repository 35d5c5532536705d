#!/bin/bash
# round-2 pass T: plan footprint (aliased scratch, low-memory mode): parity, timing A/B, sweep edge
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plan_memory.py tests/test_gpu_prologue.py tests/test_gpu_model.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r2t_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2t_pytest.log
B="--no-stock --no-eval --no-cpu-baseline --no-u8 --no-dp-parity --steps 10 --warmup 4 --report-memory"
run() {  # tag, env..., -- args
  tag=$1; shift
  env "$@" timeout 900 python bench.py $B $ARGS > gpurun_out/r2t_$tag.json 2> gpurun_out/r2t_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2t_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"], 3), "ms", round(d["value"], 1), "patches/s peak_gb", d.get("peak_gb"), d["clocks"]["sm_mhz"])
except Exception as e:
    import subprocess
    print("$tag FAILED", subprocess.run("grep -E 'Error|memory' gpurun_out/r2t_$tag.err | tail -1", shell=True, capture_output=True, text=True).stdout.strip()[:200])
PY
}
ARGS=""
run alias0 SUNET_ALIAS_SCRATCH=0
run alias1 SUNET_ALIAS_SCRATCH=1
run alias0b SUNET_ALIAS_SCRATCH=0
run alias1b SUNET_ALIAS_SCRATCH=1
run lowmem SUNET_LOW_MEM=1
ARGS="--batch 1024"
run b1024_lowmem SUNET_LOW_MEM=1 SUNET_FUSE_BNB_POOL=0
run b1024_lowmem_keeppool SUNET_LOW_MEM=1
ARGS="--batch 896"
run b896_default SUNET_LOW_MEM=0
ARGS="--batch 256 --size 512"
run s512_b256_lowmem SUNET_LOW_MEM=1 SUNET_FUSE_BNB_POOL=0
