#!/usr/bin/env python
"""Brief of one `ncu --set full` report: the handful of metrics the roofline discussion needs + top stall reasons.
usage: ncu_brief.py report.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(units, vals)))
    print('kernel:', d['Kernel Name'][1][:100])
    for k in KEYS:
        if k in d:
            print('  %-88s %-10s %s' % (k, d[k][0], d[k][1]))
    stalls = [(float(v[1]), k) for k, v in d.items() if k.startswith('smsp__average_warp') and 'issue_stalled' in k and k.endswith('_per_warp_active.pct') or
              (k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio'))]
    stalls = [(float(v[1].replace(',', '')), k) for k, v in d.items() if 'issue_stalled' in k and k.endswith('per_issue_active.ratio')]
    for v, k in sorted(stalls, reverse=True)[:8]:
        print('  stall %-70s %.2f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))


if __name__ == '__main__':
    main(sys.argv[1])
