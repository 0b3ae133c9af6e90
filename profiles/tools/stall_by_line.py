"""Attribute the warp-stall samples of an ncu report to CUDA source lines.

    python profiles/tools/stall_by_line.py <report.ncu-rep> <kernel substring> [cubin-with-lineinfo.sass]

ncu's CLI prints stall samples per SASS instruction only; nvdisasm -g on the cubin the .so carries gives the source line
of every instruction in the same order, so the two lists are zipped by position."""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
sass_file = sys.argv[3] if len(sys.argv) > 3 else '/tmp/so_mu.sass'
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
lines, cur, fn = [], None, None
for line in open(sass_file):
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*\.text\.(\S+):', line)
    if m:
        fn = m.group(1)
        continue
    if fn and kern in fn and re.match(r'\s+/\*[0-9a-f]{4,5}\*/', line):
        lines.append(cur)
print('sass instructions: ncu', len(data), 'nvdisasm', len(lines))
assert len(data) == len(lines)
isamp, iex = hdr.index('# Samples'), hdr.index('Instructions Executed')
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.defaultdict(lambda: collections.Counter())
for r, ln in zip(data, lines):
    a = agg[ln]
    a['samples'] += int(r[isamp] or 0)
    a['executed'] += int(r[iex] or 0)
    a['sass'] += 1
    for h in st:
        a[h[6:]] += int(r[hdr.index(h)] or 0)
tot = sum(a['samples'] for a in agg.values())
src = {}
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    f, l = ln if ln else ('?', 0)
    if f not in src:
        try:
            src[f] = open('/root/repo/pufferlib_b200/csrc/' + f).read().split('\n')
        except OSError:
            src[f] = []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ''
    top = sorted(((v, k) for k, v in a.items() if k not in ('samples', 'executed', 'sass')), reverse=True)[:2]
    print(f"{a['samples']:6d} {100 * a['samples'] / tot:5.1f}%  sass {a['sass']:4d} exec {a['executed']:9d}  {top[0][1]:10s} {top[1][1]:10s} {f}:{l}  {text}")
