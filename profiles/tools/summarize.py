#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python profiles/tools/summarize.py launches gpurun_out/X_launches.csv  > profiles/rNN_launches.txt
    python profiles/tools/summarize.py full     gpurun_out/X.ncu-rep       > profiles/rNN_X_full.txt
"""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__inst_executed_pipe_uniform.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg',
        'sm__cycles_elapsed.max']


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr = rows[hi]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    acc = collections.OrderedDict()
    n = 0
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        a = acc.setdefault(r[ki].split('(')[0], [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in acc.values())
    print('# ncu --metrics gpu__time_duration.sum --clock-control none : %d launches, %.1f us total' % (n, tot))
    print('# per-launch times are cold-cache and serialised: compare SHARES, not absolutes')
    print('%-72s %6s %12s %9s %7s' % ('kernel', 'count', 'total_us', 'avg_us', 'share'))
    for k, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print('%-72s %6d %12.1f %9.1f %6.1f%%' % (k[:72], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    for r in rows[2:]:
        print('== %s' % r[ki][:110])
        for w in WANT:
            for i, h in enumerate(hdr):
                if h == w:
                    print('   %-72s %s %s' % (h, r[i], units[i]))


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
