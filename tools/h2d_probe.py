"""Host->device copy rate of this box (pinned and pageable, several sizes): the ceiling of the bench's
end-to-end leg, whose timed region uploads the population once."""
import json, time, sys
import torch
out = {}
dev = torch.device('cuda', 0)
for mib in (64, 512):
    n = mib * 2 ** 20 // 8
    d = torch.empty(n, dtype=torch.float64, device=dev)
    for kind in ('pinned', 'pageable'):
        h = torch.empty(n, dtype=torch.float64, pin_memory=(kind == 'pinned'))
        h.fill_(1.0)
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out['h2d_%s_%dMiB_GBps' % (kind, mib)] = mib * 2 ** 20 / best / 1e6
        if kind == 'pinned':
            best = 1e9
            for _ in range(4):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); h.copy_(d, non_blocking=True); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out['d2h_pinned_%dMiB_GBps' % mib] = mib * 2 ** 20 / best / 1e6
import subprocess
try:
    out['pcie'] = subprocess.run(['nvidia-smi', '--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max', '--format=csv,noheader'],
                                 stdout=subprocess.PIPE, universal_newlines=True, timeout=20).stdout.strip()
except Exception as e:
    out['pcie'] = repr(e)
print(json.dumps(out))
