"""Executed warp-instructions and stall samples per SOURCE LINE of one kernel: joins the ncu SASS page (needs
--import-source on) with nvdisasm's line table of the same build.
   python tools/line_hot.py report.ncu-rep <kernel mangled-name substring> [top N]"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'f1tenth_gym_b200', 'libf110_b200.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
nvd = subprocess.run(['nvdisasm', '-g', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines = {}      # offset -> (file, line)
on, cur = False, ('?', 0)
for l in nvd:
    if l.startswith('.text.'):
        on = kern in l
        continue
    if l.startswith('.section') or (l.startswith('.') and not l.startswith('.L')) and 'text' not in l:
        pass
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+\S', l)
    if m:
        lines[int(m.group(1), 16)] = cur
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
acc = {}
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name' and kern.split('ILi')[0].replace('_ZN4f110', '')[2:8] in rows[i][1]:
        hdr = rows[i + 1]
        ia, ii, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
        j = i + 2
        base = None
        while j < len(rows) and rows[j] and rows[j][0] != 'Kernel Name':
            r = rows[j]
            if len(r) > isamp:
                try:
                    addr = int(r[ia], 16)
                    base = addr if base is None else base
                    key = lines.get(addr - base, ('?', 0))
                    a = acc.setdefault(key, [0, 0])
                    a[0] += int(r[ii] or 0)
                    a[1] += int(r[isamp] or 0)
                except ValueError:
                    pass
            j += 1
        i = j
        break
    i += 1
ti, ts = sum(a[0] for a in acc.values()), sum(a[1] for a in acc.values())
print('kernel %s: %d warp-instructions, %d samples' % (kern, ti, ts))
src_cache = {}
for (f, ln), (ie, sm) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    txt = ''
    for d in ('f1tenth_gym_b200/csrc', 'include'):
        p = os.path.join(ROOT, d, f)
        if os.path.exists(p):
            src_cache.setdefault(p, open(p).read().splitlines())
            if 0 < ln <= len(src_cache[p]):
                txt = src_cache[p][ln - 1].strip()[:90]
    print('%5.1f%% inst %5.1f%% samp  %s:%d  %s' % (100.0 * ie / max(ti, 1), 100.0 * sm / max(ts, 1), f, ln, txt))
