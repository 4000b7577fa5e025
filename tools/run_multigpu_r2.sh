#!/bin/bash
# Round-2 multi-GPU evidence (under `gpurun --gpus N`): shard parity (tests/dist_gpu_check.py), the H2D concurrency
# probe, the default bench line at N (C2 weak-scaled + c5s strong-scaled + shard_parity) and the reference arm.
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== dist check"; timeout 300 $TR tests/dist_gpu_check.py 2>&1 | grep -v "^W1\|^\*\*\*\|OMP_NUM" | tail -3 | tee gpurun_out/r2_dist_check_${N}.log
echo "== h2d"; timeout 200 $TR tools/h2d_concurrent.py 2>&1 | grep -v "^W1\|^\*\*\*\|OMP_NUM" | tee gpurun_out/r2_h2d_${N}.log | tail -40
echo "== bench N=$N"; timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/r2_bench_n${N}.err | tail -1 > gpurun_out/r2_bench_n${N}.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n${N}.json'))
print('c2', d['n_gpus'], 'ms %.4f' % d['ms_per_step'], 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['e2e'].get('breakdown'), 'parity', d.get('shard_parity'), d['clocks'])
for k, v in d.get('also', {}).items():
    print(k, v.get('ms_per_step'), '%.4g' % v.get('value', 0), v.get('roofline', {}).get('frac'), v.get('clocks'), v.get('error'))
PY
tail -c 600 gpurun_out/r2_bench_n${N}.err
echo "== reference N=$N"; timeout 600 $TR bench.py --impl reference --gpus $N --steps 5 --warmup 1 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('reference', d['value'], d['cpu_baseline']['cores'], d['config'])"
