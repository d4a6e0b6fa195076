"""Static size of a kernel's code by source line (kernel experiments: instruction-cache footprint).

    python scripts/sass_static_lines.py aslsd_b200/_lib/obj_libaslsd_b200/exact.o lag_moment_warp_kernel [n_top]

Disassembles the object's sm_100a cubin with nvdisasm -g (needs -lineinfo) and counts the SASS
instructions of the named function per source line and per source function of lag_quadrature.cuh."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj, kern = sys.argv[1], sys.argv[2]
n_top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
sass = subprocess.run(['nvdisasm', '-g', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split('\n')
inside = False
cur = ('?', 0)
per_line = collections.Counter()
n = 0
sections = collections.Counter()
sec = None
for line in sass:
    m = re.match(r'\s*\.section\s+(\S+),', line)
    if m:
        sec = m.group(1)
        inside = kern in sec and sec.startswith('.text.')
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', line):
        if sec and sec.startswith('.text.'):
            sections[sec] += 1
        if inside:
            per_line[cur] += 1
            n += 1
print('%s: %d instructions (%.1f KB)' % (kern, n, n * 16 / 1024.))
for s, c in sections.most_common():
    if 'lqw' in s or kern in s:
        print('  section %-90s %6d instructions' % (s[:90], c))
src = open(os.path.join(ROOT, 'aslsd_b200', 'csrc', 'lag_quadrature.cuh')).read().split('\n')
starts = []
for i, line in enumerate(src, 1):
    if line.startswith(' '):
        continue
    m = re.match(r'^(?:__device__ (?:__forceinline__|__noinline__) \S+ |lag_\w+\()?(\w+)\(', line)
    if m and ('__device__' in line or line.startswith('lag_')):
        starts.append((i, m.group(1)))
starts.sort()
per_fn = collections.Counter()
for (f, ln), c in per_line.items():
    if f != 'lag_quadrature.cuh':
        per_fn[f] += c
        continue
    name = '?'
    for s, nm in starts:
        if s <= ln:
            name = nm
    per_fn[name] += c
for name, c in per_fn.most_common():
    print('%-40s %6d  %5.1f %%' % (name, c, 100. * c / n))
print('hottest lines by static count:')
for (f, ln), c in per_line.most_common(n_top):
    text = src[ln - 1].strip()[:100] if f == 'lag_quadrature.cuh' and ln <= len(src) else ''
    print('%-22s %5d %5d  %s' % (f, ln, c, text))
