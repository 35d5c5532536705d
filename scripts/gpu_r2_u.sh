#!/bin/bash
# round-2 pass U: extended scratch aliasing + single level-1 dY buffer as default: parity, A/B, batch 1024 out of the box
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plan_memory.py tests/test_gpu_prologue.py tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_resume_val_ensemble.py -x -q > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2u_pytest.log
B="--no-stock --no-eval --no-cpu-baseline --no-u8 --no-dp-parity --steps 10 --warmup 4 --report-memory"
run() {
  tag=$1; shift
  env "$@" timeout 900 python bench.py $B $ARGS > gpurun_out/r2u_$tag.json 2> gpurun_out/r2u_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2u_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"], 3), "ms", round(d["value"], 1), "patches/s peak_gb", d.get("peak_gb"), d["clocks"]["sm_mhz"])
except Exception as e:
    import subprocess
    print("$tag FAILED", subprocess.run("grep -E 'Error|memory' gpurun_out/r2u_$tag.err | tail -1", shell=True, capture_output=True, text=True).stdout.strip()[:200])
PY
}
ARGS=""
run two_dy SUNET_LOW_MEM=0
run one_dy SUNET_LOW_MEM=1
run two_dy_b SUNET_LOW_MEM=0
run one_dy_b SUNET_LOW_MEM=1
ARGS="--batch 16"
run b16_two_dy SUNET_LOW_MEM=0
run b16_one_dy SUNET_LOW_MEM=1
run b16_two_dy_b SUNET_LOW_MEM=0
run b16_one_dy_b SUNET_LOW_MEM=1
ARGS="--batch 1024"
run b1024_default SUNET_LOW_MEM=1
ARGS="--batch 256 --size 512 --non-selective"
run s512_b256_unet_default SUNET_LOW_MEM=1
