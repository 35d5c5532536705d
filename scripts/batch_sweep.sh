#!/bin/bash
# BASELINE configs[4]: UNet_B and SUNet_B batch-size sweep at 256^2 and 512^2 up to the out-of-memory edge.
# usage: scripts/batch_sweep.sh <out file>
set -u
out=${1:-gpurun_out/batch_sweep.txt}
echo "# model  size  batch  ms/step  patches/s  256^2-equivalent patches/s  peak GB allocated  (or the failure)" > $out
B="--no-stock --no-eval --no-cpu-baseline --no-u8 --steps 5 --warmup 3"
for model in SUNet_B UNet_B; do
  flag=""; [ $model = UNet_B ] && flag="--non-selective"
  for cfg in "256 32" "256 64" "256 128" "256 256" "256 512" "256 768" "256 1024" "256 1152" "512 32" "512 64" "512 128" "512 192" "512 256" "512 288"; do
    set -- $cfg
    python - "$model" "$1" "$2" $flag <<'PY' >> $out 2>/dev/null
import json, subprocess, sys
model, size, batch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cmd = [sys.executable, "bench.py", "--size", str(size), "--batch", str(batch), "--no-stock", "--no-eval", "--no-cpu-baseline",
       "--no-u8", "--steps", "5", "--warmup", "3", "--report-memory"] + sys.argv[4:]
r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
line = [l for l in r.stdout.splitlines() if l.startswith("{")]
if r.returncode == 0 and line:
    d = json.loads(line[-1])
    eq = d["value"] * (size / 256) ** 2
    print(f"{model:8s} {size:4d} {batch:5d} {d['ms_per_step']:9.2f} {d['value']:10.1f} {eq:10.1f} {d.get('peak_gb', float('nan')):8.1f}")
else:
    err = (r.stderr or r.stdout).strip().splitlines()
    msg = next((l for l in reversed(err) if "Error" in l or "error" in l or "memory" in l), err[-1] if err else "?")
    print(f"{model:8s} {size:4d} {batch:5d}   FAILED: {msg[:140]}")
PY
  done
done
cat $out
