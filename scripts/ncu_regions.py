"""Warp instructions and samples of an .ncu-rep per function of the warp-shaped moment kernel
(needs -lineinfo and --import-source on).

    python scripts/ncu_regions.py rep.ncu-rep [event pairs binned]

Lines of lag_quadrature.cuh are assigned to the device function they lie in (the kernel body is split into
its set-up, the tile prologue and the mark / chunk loop); lines of other files to the file.  With the number
of event pairs the launch binned, instructions are also given per step of 32 pairs."""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
pairs = float(sys.argv[2]) if len(sys.argv) > 2 else None
src = open(os.path.join(ROOT, 'aslsd_b200', 'csrc', 'lag_quadrature.cuh')).read().split('\n')
starts = []   # (line, name)
for i, line in enumerate(src, 1):
    m = re.match(r'^(?:__device__ __forceinline__ \S+ |lag_\w+\()?(\w+)\(', line) if not line.startswith(' ') else None
    if m and ('__device__' in line or line.startswith('lag_')):
        starts.append((i, m.group(1)))
    if 'for (int pass = 0; pass < n_pass; ++pass)' in line:
        starts.append((i, 'warp kernel: pass / tile prologue'))
    if '// ---- b. steps of 32 pairs, LQW_PC pair slots at a time' in line:
        starts.append((i, 'warp kernel: marks, chunk loop, final drain, write-out'))
starts.sort()


def region(f, ln):
    if f != 'lag_quadrature.cuh':
        return 'inlined from ' + f
    name = 'file header'
    for s, n in starts:
        if s <= ln:
            name = n
    return name


out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
cur_file, hdr = None, None
inst, samp = collections.Counter(), collections.Counter()
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    elif len(r) > 6 and r[0] == 'Line No':
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 6 and r[2] == '-':
        try:
            ln, s, i = int(r[0]), int(r[hdr['# Samples']] or 0), int(r[hdr['Instructions Executed']] or 0)
        except ValueError:
            continue
        inst[region(cur_file, ln)] += i
        samp[region(cur_file, ln)] += s
ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print('warp instructions (source view) %d, samples %d' % (ti, ts))
for k, v in inst.most_common():
    per = '  %6.1f per 32 pairs' % (v / (pairs / 32.)) if pairs else ''
    print('%-62s inst %5.1f %%%s   samples %5.1f %%' % (k, 100. * v / ti, per, 100. * samp[k] / ts))
if pairs:
    print('total per 32 pairs: %.1f' % (ti / (pairs / 32.)))
