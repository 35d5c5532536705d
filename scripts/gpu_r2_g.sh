#!/bin/bash
# round-2 multi-GPU pass: bench.py under torchrun (dp_parity block + scaling point)
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err
echo "rc=$?"
tail -3 gpurun_out/r2g_bench_n$N.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r2g_bench_n$N.json') if l.startswith('{')][-1]
print('N=$N', round(d['ms_per_step'],3),'ms/step', round(d['value'],1),'patches/s e2e', round(d['e2e']['value'],1), d['clocks'])
print('dp_parity', json.dumps(d.get('dp_parity')))
print('eval', d['eval']['value'], d['eval']['pixels_counted'], d['eval']['pixels_expected'])
"
