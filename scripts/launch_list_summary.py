"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file ...).

    python scripts/launch_list_summary.py gpurun_out/launch_list.csv [label]
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else path
rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
tot, cnt, moments = collections.Counter(), collections.Counter(), []
for r in rows:
    name = re.sub(r'\(.*', '', r[4])
    us = float(r[-1].replace(',', '')) * {'ns': 1e-3, 'us': 1., 'ms': 1e3}.get(r[-2], 1e-3)
    tot[name] += us
    cnt[name] += 1
    if 'lag_moment' in name:
        moments.append(us)
print('ncu launch list of `python bench.py --steps 2 --warmup 3 --no_cpu_baseline` (%s): per kernel, all launches of the '
      'run (serialised under ncu: the two lanes of a step do not overlap here)' % label)
print('%-70s %6s %12s' % ('kernel', 'count', 'total us'))
for name, us in tot.most_common():
    print('%-70s %6d %12.1f' % (name[:70], cnt[name], us))
full = [m for m in moments if m > 0.5 * max(moments)] if moments else []
if full:
    print('moment-kernel launches of a full step: %d, mean %.1f us' % (len(full), sum(full) / len(full)))
