"""dram traffic of the dominant kernel from an ``ncu --set full`` capture -> profiles/lag_moment_kernel_dram.json,
the file bench.py reads ``roofline.traffic`` from.

    python scripts/ncu_traffic.py gpurun_out/r2_moment.ncu-rep [label]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(rep)
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
best = None
for vals in rows[2:]:
    if 'lag_moment_' not in vals[ix['Kernel Name']]:
        continue
    dur = float(vals[ix['gpu__time_duration.sum']].replace(',', ''))
    if best is None or dur > best[0]:
        best = (dur, vals)
if best is None:
    sys.exit('no lag_moment_* launch in ' + rep)
vals = best[1]


def num(name):
    return float(vals[ix[name]].replace(',', '')) * scale.get(units[ix[name]], 1.)


out = {'kernel': vals[ix['Kernel Name']][:60], 'dram_bytes_read': num('dram__bytes_read.sum'),
       'dram_bytes_write': num('dram__bytes_write.sum'),
       'duration_ms_under_ncu': float(vals[ix['gpu__time_duration.sum']].replace(',', '')) *
       {'ns': 1e-6, 'us': 1e-3, 'ms': 1., 's': 1e3}.get(units[ix['gpu__time_duration.sum']], 1.),
       'grid_size': vals[ix['launch__grid_size']], 'capture': label,
       'how': "ncu --set full --clock-control none, one launch of the C5 step (bench.py's step, or scripts/profile_one.py 1.1e6: same stream, model, evaluation point and batched call)"}
with open(os.path.join(ROOT, 'profiles', 'lag_moment_kernel_dram.json'), 'w') as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
