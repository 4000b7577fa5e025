#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_c2_launches.csv
  python tools/summarize_ncu.py full     gpurun_out/prof.ncu-rep    profiles/r1_c2_tma_full.csv
  python tools/summarize_ncu.py source   gpurun_out/prof.ncu-rep    profiles/r1_c2_tma_sass_mix.csv
"""
import collections
import csv
import subprocess
import sys

KEEP = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def ncu_csv(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, universal_newlines=True).stdout
    return list(csv.reader(out.splitlines()))


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    ix = dict((h, i) for i, h in enumerate(rows[0]))
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ix['Kernel Name']].split('(')[0].replace('void ', '')
        agg.setdefault(name, []).append(float(r[ix['Metric Value']]) / 1e3)  # ns -> us
    tot = sum(sum(v) for v in agg.values())
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'launches', 'total_us', 'mean_us', 'min_us', 'max_us', 'share_of_captured_gpu_time'])
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            w.writerow([k, len(v), '%.1f' % sum(v), '%.2f' % (sum(v) / len(v)), '%.2f' % min(v), '%.2f' % max(v),
                        '%.4f' % (sum(v) / tot)])


def full(rep, dst):
    rows = ncu_csv(rep, 'raw')
    hdr, units = rows[0], rows[1]
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['launch', 'metric', 'unit', 'value'])
        for li, vals in enumerate(rows[2:]):
            for i, h in enumerate(hdr):
                if h in KEEP:
                    w.writerow([li, h, units[i], vals[i]])


def source(rep, dst):
    rows = ncu_csv(rep, 'source')
    hdr = rows[1]
    ix = dict((h, i) for i, h in enumerate(hdr))
    byop, stall, tot = collections.Counter(), collections.Counter(), 0
    stalls = [h for h in hdr if h.startswith('stall_') and '(' not in h]
    for r in rows[2:]:
        try:
            n = int(r[ix['Instructions Executed']])
        except (ValueError, IndexError):
            continue
        t = r[ix['Source']].split()
        op = (t[1] if t and t[0].startswith('@') else (t[0] if t else '')).split('.')[0]
        byop[op] += n
        tot += n
        for k in stalls:
            try:
                stall[k] += int(r[ix[k]])
            except ValueError:
                pass
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kind', 'name', 'count', 'share'])
        for op, n in byop.most_common(40):
            w.writerow(['sass_opcode_warp_instructions', op, n, '%.4f' % (n / tot)])
        ts = sum(stall.values()) or 1
        for k, n in stall.most_common():
            w.writerow(['warp_stall_samples', k, n, '%.4f' % (n / ts)])


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'source': source}[sys.argv[1]](sys.argv[2], sys.argv[3])
