"""Per-SASS-instruction executed counts and stall samples from an ncu report (needs --import-source on):
   python tools/sass_hot.py report.ncu-rep [top N]   ->  instruction totals by address range, hottest instructions."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
ci = {n: hdr.index(n) for n in ('Address', 'Source', '# Samples', 'Instructions Executed', 'Avg. Threads Executed')}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        data.append((r[ci['Address']], r[ci['Source']], int(r[ci['# Samples']] or 0), int(r[ci['Instructions Executed']] or 0),
                     float(r[ci['Avg. Threads Executed']] or 0)))
    except ValueError:
        pass
tot_i = sum(d[3] for d in data)
tot_s = sum(d[2] for d in data)
print('total warp-instructions %d, samples %d, %d SASS lines' % (tot_i, tot_s, len(data)))
print('--- all instructions in address order: idx addr  inst%%  samp%%  thr  sass')
for k, d in enumerate(data):
    if d[3] * 1000 >= tot_i or d[2] * 1000 >= tot_s:
        print('%4d %s %5.2f %5.2f %4.1f  %s' % (k, d[0][-5:], 100.0 * d[3] / tot_i, 100.0 * d[2] / max(tot_s, 1), d[4], d[1][:90]))
