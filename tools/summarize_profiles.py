"""Turns the ncu captures brought back under gpurun_out/ into the small, tracked summaries
under profiles/ (launch list -> per-kernel shares; `--set full` reports -> key metrics)."""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles')

METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size',
    'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
]


def launches(csv_path, command, note):
    rows = []
    with open(csv_path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            rows.append((r['Kernel Name'], float(r['Metric Value']) / 1e3))
    tot = sum(t for _, t in rows)
    agg = OrderedDict()
    for k, t in rows:
        k = k.split('(')[0][:100]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    ks = sorted(agg.items(), key=lambda kv: -kv[1][1])
    return dict(command=command, note=note, total_us=round(tot, 1), launches=len(rows),
                kernels=[dict(kernel=k, launches=n, total_us=round(t, 1), share=round(t / tot, 4))
                         for k, (n, t) in ks])


def full(rep_path):
    if rep_path.endswith('.csv'):      # already exported on the GPU box (`ncu -i rep --page raw --csv`)
        out = open(rep_path).read()
    else:
        out = subprocess.run(['ncu', '-i', rep_path, '--page', 'raw', '--csv'], capture_output=True,
                             text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = OrderedDict(kernel=r[hdr.index('Kernel Name')][:110])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                d[m] = ('%s %s' % (r[i], units[i])).strip()
        res.append(d)
    return res


if __name__ == '__main__':
    g = os.path.join(ROOT, 'gpurun_out')
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
    jobs = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
    if 'launches' in jobs:
        j = jobs['launches']
        s = launches(os.path.join(g, j['csv']), j['command'], j.get('note', ''))
        json.dump(s, open(os.path.join(OUT, '%s_launches_summary.json' % tag), 'w'), indent=1)
        # keep the raw launch list too (small)
        with open(os.path.join(g, j['csv'])) as f, open(os.path.join(OUT, '%s_launches.csv' % tag), 'w') as o:
            o.write(f.read())
    if 'full' in jobs:
        allk = []
        for rep in jobs['full']:
            allk.extend(full(os.path.join(g, rep)))
        json.dump(allk, open(os.path.join(OUT, '%s_ncu_full_summary.json' % tag), 'w'), indent=1)
    print('ok')
