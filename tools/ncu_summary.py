#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
  python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep       > profiles/rNN_<kernel>_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

RAW_KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
            'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
            'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[row['Metric Unit']]
        e = agg.setdefault(k, [0, 0.0])
        e[0] += 1
        e[1] += v
        n += 1
    tot = sum(e[1] for e in agg.values())
    print('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare SHARES)')
    print('# %d launches, %.2f ms total device time' % (n, tot))
    print('%-72s %6s %12s %7s' % ('kernel', 'calls', 'ms', 'share'))
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if ms / tot < 0.0005:
            continue
        print('%-72s %6d %12.3f %6.1f%%' % (k[:72], c, ms, 100 * ms / tot))


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print('# ncu --set full --clock-control none: selected raw metrics per captured launch (%s)' % path)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('--- %s  grid %s block %s' % (d.get('Kernel Name', '?')[:100], d.get('Grid Size'), d.get('Block Size')))
        for k in RAW_KEYS:
            if k in d:
                print('  %-86s %s %s' % (k, d[k], units[hdr.index(k)]))


def rawcsv(path):
    """Same summary from a `ncu -i rep --page raw --csv` export made on the GPU box (the .ncu-rep itself is too big to bring back)."""
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    extra = ['l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
             'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
             'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
             'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
    print('# ncu --set full --clock-control none: selected raw metrics per captured launch (%s)' % path)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('--- %s  grid %s block %s' % (d.get('Kernel Name', '?')[:110], d.get('Grid Size'), d.get('Block Size')))
        for k in RAW_KEYS + extra:
            if k in d:
                print('  %-86s %s %s' % (k, d[k], units[hdr.index(k)]))


if __name__ == '__main__':
    {'launches': launches, 'raw': raw, 'rawcsv': rawcsv}[sys.argv[1]](sys.argv[2])
