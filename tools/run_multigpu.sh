#!/bin/bash
# Multi-GPU evidence run (under `gpurun --gpus N`): parity of every shard with the oracle's emulation
# (tests/dist_gpu_check.py) and the weak-scaling bench lines for C2 and C5.   usage: run_multigpu.sh N
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== dist check exchange=peer"; timeout 200 $TR tests/dist_gpu_check.py 2>&1 | grep -v "^W1\|^\*\*\*\|OMP_NUM" | tail -4 | tee gpurun_out/dist_check_${N}_peer.log
for w in c2 c5; do for ex in peer; do
  echo "== bench $w N=$N exchange=$ex"; timeout 200 $TR bench.py --gpus $N --workload $w --exchange $ex --no-cpu-baseline 2>gpurun_out/bench_${w}_n${N}_${ex}.err | tail -1 > gpurun_out/bench_${w}_n${N}_${ex}.json; python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_n${N}_${ex}.json')); print(d['n_gpus'], d['ms_per_step'], d['config']['burst_ms_per_step'], '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], d['gpu_launches'], d['clocks'])"
done; done
