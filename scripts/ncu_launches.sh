#!/bin/bash
# Launch list of one bench run (B200_PROFILING.md recipe): plain run first, then the gpu__time_duration pass.
# usage: scripts/ncu_launches.sh <tag> [extra bench.py args]     -> gpurun_out/launches_<tag>.csv
set -u
tag=${1:-r01}
shift
extra="$@"
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline $extra > gpurun_out/plain_launches_$tag.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline $extra \
    > gpurun_out/ncu_launches_$tag.log 2>&1
echo "ncu rc=$?"
