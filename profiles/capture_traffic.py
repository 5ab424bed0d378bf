#!/usr/bin/env python
"""DRAM bytes of ONE launch of each workload's dominant kernel (`roofline.traffic` of bench.py), from ncu:

    python profiles/capture_traffic.py OUTDIR [WORKLOAD ...]     (on a GPU box; writes OUTDIR/traffic.json + the raw CSVs;
                                                                  workloads given: only those)

For every workload: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` over a 3-step bench run,
the launches of the dominant kernel are filtered by name and the LONGEST one (layer 1 on the (x1, x2) pair) is taken."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORK = [('reddit', 16384, 'gather_reduce_kernel'), ('pokec-mean', 32768, 'gather_reduce_kernel'), ('big10m', 16384, 'gather_reduce_kernel'),
        ('pokec-maxpool', 16384, 'linear_pool_ws_umma_kernel'), ('plaw2m-attention', 16384, 'attention_fused_kernel')]


def main(out):
    os.makedirs(out, exist_ok=True)
    res = {'_note': 'dram__bytes_read.sum + dram__bytes_write.sum of the LONGEST launch of the dominant kernel of each workload '
                    '(layer 1 on the (x1, x2) pair), ncu --metrics pass over `bench.py --workload W --batch B --steps 3`; '
                    'made by profiles/capture_traffic.py'}
    only = sys.argv[2:]
    for wl, B, kern in WORK:
        if only and wl not in only:
            continue
        log = os.path.join(out, 'traffic_%s.csv' % wl)
        cmd = ['ncu', '--metrics', 'dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum', '--clock-control', 'none',
               '-k', 'regex:' + kern, '-c', '40', '--csv', '--log-file', log,
               sys.executable, os.path.join(ROOT, 'bench.py'), '--workload', wl, '--batch', str(B), '--steps', '3', '--warmup', '3',
               '--no-train', '--no-cpu-baseline', '--no-ahead']
        subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
        rows = {}
        lines = [l for l in open(log) if not l.startswith('==')]
        for r in csv.DictReader(lines):
            v = float(r['Metric Value'].replace(',', ''))
            u = r['Metric Unit']
            if r['Metric Name'].startswith('dram__bytes'):
                v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
            else:
                v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(u, 1)
            rows.setdefault(r['ID'], {})[r['Metric Name']] = v
        if not rows:
            continue
        best = max(rows.values(), key=lambda m: m.get('gpu__time_duration.sum', 0))
        res['%s:%d' % (wl, B)] = int(best['dram__bytes_read.sum'] + best['dram__bytes_write.sum'])
        res['_us_%s:%d' % (wl, B)] = best['gpu__time_duration.sum']
    json.dump(res, open(os.path.join(out, 'traffic.json'), 'w'), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/traffic')
