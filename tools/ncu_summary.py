#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers DESIGN.md/profiles quote."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[hdr.index('Kernel Name')][:70])
        for k in KEYS:
            if k in hdr:
                print('   %-62s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = [(float(r[i].replace(',', '') or 0), hdr[i]) for i in range(len(hdr)) if 'smsp__average_warps_issue_stalled' in hdr[i] and hdr[i].endswith('per_issue_active.ratio')]
        for v, k in sorted(st, reverse=True)[:7]:
            print('   stall %-56s %.2f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))


if __name__ == '__main__':
    main(sys.argv[1])
