"""Samples and executed instructions per CUDA source line of an .ncu-rep (needs -lineinfo and --import-source on).

    python scripts/ncu_lines.py rep.ncu-rep [n_top]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, data = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    elif len(r) > 6 and r[0] == 'Line No':
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 6 and r[2] == '-':      # a CUDA line (its SASS rows carry an address)
        try:
            data.append((cur_file, int(r[0]), r[1], int(r[hdr['# Samples']] or 0), int(r[hdr['Instructions Executed']] or 0),
                         {k[6:]: int(r[i] or 0) for k, i in hdr.items() if k.startswith('stall_') and 'Not Issued' not in k}))
        except ValueError:
            pass
ts = sum(d[3] for d in data) or 1
ti = sum(d[4] for d in data) or 1
print('total samples %d, warp instructions %d' % (ts, ti))
for f, ln, src, s, i, st in sorted(data, key=lambda d: -d[3])[:n_top]:
    top = ', '.join('%s %.0f%%' % (k, 100 * v / max(s, 1)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print('%5.2f%% samp %5.2f%% inst  %s:%d  %s   [%s]' % (100 * s / ts, 100 * i / ti, f, ln, src.strip()[:70], top))
