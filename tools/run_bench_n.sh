#!/bin/bash
# the default bench line under torchrun on N GPUs (under `gpurun --gpus N`)
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/r2c_bench_n${N}.err | tail -1 > gpurun_out/r2c_bench_n${N}.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2c_bench_n${N}.json'))
print('c2', d['n_gpus'], 'ms %.4f' % d['ms_per_step'], 'burst', d['details'].get('burst_ms_per_step'), 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'parity', d.get('shard_parity'), d['clocks'])
for k, v in d.get('also', {}).items():
    print(k, v.get('ms_per_step'), '%.4g' % v.get('value', 0), v.get('roofline', {}).get('frac'), v.get('error'))
PY
