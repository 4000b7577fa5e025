#!/usr/bin/env python
"""Host->device upload rate with 1..N ranks uploading at once (torchrun, one rank per GPU): the ceiling of the bench's
end-to-end leg at N GPUs.  Prints the host topology first (NUMA nodes, GPU affinities), then for each group size the
slowest rank's time for 512 MiB from pinned memory."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=30).stdout
    except Exception as e:
        return repr(e)


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    if rank == 0:
        print(sh('nvidia-smi topo -m'))
        print(sh('lscpu | grep -i -E "numa|socket|model name|^cpu\\(s\\)"'))
        print(sh('ls /sys/devices/system/node/ | head; cat /sys/devices/system/node/node*/cpulist 2>/dev/null'))
        print('affinity', sorted(os.sched_getaffinity(0)), flush=True)
    mib = 512
    n = mib * 2 ** 20 // 8
    d = torch.empty(n, dtype=torch.float64, device='cuda')
    h = torch.empty(n, dtype=torch.float64, pin_memory=True)
    h.fill_(1.0)
    out = {}
    for group in [1, 2, 4, 8]:
        if group > world:
            break
        for rep in range(3):
            torch.cuda.synchronize()
            dist.barrier()
            ms = 0.0
            if rank < group:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                d.copy_(h, non_blocking=True)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
            t = torch.tensor([ms], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out['%d ranks' % group] = min(out.get('%d ranks' % group, 1e9), float(t.item()))
    if rank == 0:
        print(json.dumps({k: {'ms_slowest_rank': v, 'GBps_per_rank': mib * 2 ** 20 / v / 1e6, 'GBps_total': int(k.split()[0]) * mib * 2 ** 20 / v / 1e6}
                          for k, v in out.items()}), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
