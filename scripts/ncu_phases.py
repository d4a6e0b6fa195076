import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, data = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur_file = r[1].split('/')[-1]
    elif len(r) > 6 and r[0] == 'Line No': hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 6 and r[2] == '-':
        try: data.append((cur_file, int(r[0]), int(r[hdr['# Samples']] or 0), int(r[hdr['Instructions Executed']] or 0)))
        except ValueError: pass
regions = eval(sys.argv[2])
ts = sum(d[2] for d in data); ti = sum(d[3] for d in data)
acc = {name: [0, 0] for name, _ in regions}; acc['other'] = [0, 0]
for f, ln, s, i in data:
    for name, test in regions:
        if test(f, ln):
            acc[name][0] += s; acc[name][1] += i; break
    else:
        acc['other'][0] += s; acc['other'][1] += i
for k, (s, i) in acc.items(): print('%-14s samples %5.1f%%  instr %5.1f%%  (%.1f instr / 32 pairs)' % (k, 100*s/ts, 100*i/ti, i / (1397291792/32)))
