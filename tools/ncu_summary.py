"""Summarise ncu outputs: python tools/ncu_summary.py launches FILE.csv | raw FILE.ncu-rep"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']


def launches(path, per_launch=False):
    rows = list(csv.reader(l for l in open(path) if not l.startswith('==')))
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    seq = []
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r'\(.*', '', r[ki]).replace('void ', '')
        v = float(r[vi].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(r[ui], 1e-6)
        agg[name][0] += 1
        agg[name][1] += v
        seq.append((name, v))
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':48s} {'n':>5s} {'total ms':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:48s} {v[0]:5d} {v[1]:10.3f} {v[1] / tot * 100:6.1f}%")
    if per_launch:
        for name, v in seq:
            print(f"  {name:46s} {v:9.4f} ms")


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for k in ['Kernel Name'] + KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:82s} {units[i]:10s} " + "  ".join(r[i][:40] for r in rows[2:]))


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], per_launch=len(sys.argv) > 3)
    else:
        raw(sys.argv[2])
