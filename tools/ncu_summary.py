"""Key metrics of an ncu report as text (one column per captured launch) + the per-launch DRAM traffic as JSON.
   python tools/ncu_summary.py report.ncu-rep [kernel-substring] [--json out.json --workload cfg2]"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex_op_read.sum',
    'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
    'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'smsp__warps_eligible.avg.per_cycle_active',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
    'smsp__average_warp_latency_per_inst_issued.ratio',
]


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith('--') else ''
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kcol = hdr.index('Kernel Name')
    data = [r for r in data if sub in r[kcol]]
    print('# %s: %d launches of %s' % (rep.split('/')[-1], len(data), sorted(set(r[kcol] for r in data))))
    vals = {}
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            vals[m] = [r[i] for r in data]
            print('%-86s %s  %s' % (m, ' '.join('%-14s' % v for v in vals[m]), units[i]))
    if '--json' in sys.argv:
        path = sys.argv[sys.argv.index('--json') + 1]
        wl = sys.argv[sys.argv.index('--workload') + 1] if '--workload' in sys.argv else 'cfg2'

        def to_bytes(m):
            i = hdr.index(m)
            scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[i]]
            return [float(r[i].replace(',', '')) * scale for r in data]
        rd, wr = to_bytes('dram__bytes_read.sum'), to_bytes('dram__bytes_write.sum')
        per = sum(a + b for a, b in zip(rd, wr)) / len(rd)
        try:
            cur = json.load(open(path))
        except Exception:
            cur = {}
        cur[wl] = {'kernel': sorted(set(r[kcol] for r in data))[0], 'launches': len(rd),
                   'dram_bytes_per_launch': per, 'dram_read_bytes_per_launch': sum(rd) / len(rd),
                   'dram_write_bytes_per_launch': sum(wr) / len(wr), 'source': rep.split('/')[-1]}
        json.dump(cur, open(path, 'w'), indent=1)


if __name__ == '__main__':
    main()
