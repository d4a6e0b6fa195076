"""Group the SASS lines of an .ncu-rep by execution count (loop nest level) and list one group.

    python scripts/ncu_groups.py rep.ncu-rep [exec_count_to_list]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix['Instructions Executed']]) for r in data)
tots = sum(int(r[ix['# Samples']]) for r in data)
groups = collections.defaultdict(lambda: [0, 0, 0])
for r in data:
    c = int(r[ix['Instructions Executed']])
    key = float('%.2g' % c)
    g = groups[key]
    g[0] += c
    g[1] += int(r[ix['# Samples']])
    g[2] += 1
for k in sorted(groups, key=lambda k: -groups[k][0])[:14]:
    g = groups[k]
    print('exec~%.2e: n_instr=%4d  inst=%5.1f%%  samples=%5.1f%%' % (k, g[2], 100 * g[0] / tot, 100 * g[1] / tots))
if len(sys.argv) > 2:
    want = float(sys.argv[2])
    for r in data:
        c = int(r[ix['Instructions Executed']])
        if abs(c - want) < 0.06 * want:
            print('%6.2f%% %s' % (100 * int(r[ix['# Samples']]) / tots, r[ix['Source']][:90]))
