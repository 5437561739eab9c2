#!/bin/bash
# Turns the raw captures of tools/capture_profiles.sh (gpurun_out/cap_*) into the committed summaries under profiles/<round>/
# and profiles/raymarch_ncu_summary.json.   bash tools/refresh_profiles.sh r2
R=${1:-r2}
P=profiles/$R
mkdir -p $P
python - <<'PY'
import csv, io, json
path = 'profiles/raymarch_ncu_summary.json'
try:
    cur = json.load(open(path))
except Exception:
    cur = {}
for b in (270, 540, 1080, 2160):
    try:
        txt = [l for l in open('gpurun_out/cap_march_cfg5_%d.csv' % b) if l.startswith('"')]
    except OSError:
        continue
    rows = list(csv.DictReader(io.StringIO(''.join(txt))))
    rd = [float(r['Metric Value']) for r in rows if r['Metric Name'] == 'dram__bytes_read.sum']
    wr = [float(r['Metric Value']) for r in rows if r['Metric Name'] == 'dram__bytes_write.sum']
    cur['cfg5_%d' % b] = {'kernel': rows[0]['Kernel Name'], 'launches': len(rd), 'dram_bytes_per_launch': (sum(rd) + sum(wr)) / len(rd),
                          'dram_read_bytes_per_launch': sum(rd) / len(rd), 'dram_write_bytes_per_launch': sum(wr) / len(wr),
                          'source': 'cap_march_cfg5_%d.csv' % b}
json.dump(cur, open(path, 'w'), indent=1)
PY
for w in cfg3 cfg2 cfg2x2; do
  python tools/ncu_summary.py gpurun_out/cap_march_$w.ncu-rep k_march_lean --json profiles/raymarch_ncu_summary.json --workload $w > $P/march_lean_${w}_ncu.txt
done
cp gpurun_out/cap_launches.csv $P/launches_bench_cfg3.csv
python tools/launch_summary.py gpurun_out/cap_launches.csv > $P/launches_bench_cfg3_summary.txt
K=$(grep -o 'k_march_lean<[^>]*>' $P/march_lean_cfg3_ncu.txt | head -1)
echo "march kernel at cfg3: $K"
python tools/line_hot.py gpurun_out/cap_march_cfg3.ncu-rep _ZN4f11012k_march_leanILi0ELb0ELb0ELb1ELb0ELi 30 > $P/march_lean_cfg3_line_hot.txt
python tools/ncu_summary.py gpurun_out/cap_dyn_tail_cfg3.ncu-rep k_dynamics > $P/dynamics_cfg3_ncu.txt
python tools/ncu_summary.py gpurun_out/cap_dyn_tail_cfg3.ncu-rep k_tail > $P/tail_cfg3_ncu.txt
python tools/line_hot.py gpurun_out/cap_dyn_tail_cfg3.ncu-rep _ZN4f11010k_dynamicsE 25 > $P/dynamics_cfg3_line_hot.txt
