#!/usr/bin/env python
"""Per-source-line totals (warp instructions, stall samples) of one kernel from an .ncu-rep that was
captured with --import-source on and compiled with -lineinfo.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True).stdout
    rows = list(csv.reader(out.splitlines()))
    path = ''
    hdr = None
    lines = []
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            path = r[1].split('/')[-1]
            continue
        if len(r) > 10 and r[0] == 'Line No':
            hdr = dict((h, i) for i, h in enumerate(r))
            continue
        if hdr is None or len(r) < 10 or not r[0]:
            continue
        try:
            inst = int(r[hdr['Instructions Executed']])
            samp = int(r[hdr['# Samples']])
        except ValueError:
            continue
        lines.append((inst, samp, path, r[0], r[1].strip()[:110]))
    tot_i = sum(l[0] for l in lines) or 1
    tot_s = sum(l[1] for l in lines) or 1
    print('total warp instructions %d, samples %d' % (tot_i, tot_s))
    print('-- by instructions')
    for inst, samp, p, ln, src in sorted(lines, reverse=True)[:top]:
        print('%5.1f%% inst %5.1f%% samp  %s:%s  %s' % (100. * inst / tot_i, 100. * samp / tot_s, p, ln, src))
    print('-- by stall samples')
    for inst, samp, p, ln, src in sorted(lines, key=lambda l: -l[1])[:top // 2]:
        print('%5.1f%% inst %5.1f%% samp  %s:%s  %s' % (100. * inst / tot_i, 100. * samp / tot_s, p, ln, src))


if __name__ == '__main__':
    main()
