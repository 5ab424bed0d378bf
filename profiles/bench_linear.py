#!/usr/bin/env python
"""Micro-benchmark of the layer projection [fc_x(table[ids]) | fc_neib(M)] + relu (linear_ws_umma.cu) on a layer-1 shape.

    MODE=x3|tf32|bf16  D=64  ROWS=1632803  N=851968  python profiles/bench_linear.py
    (GSAGE_B200_LIB=.../libgsage_b200_timing.so prints per-role cycle counters of CTA 0)"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_graphsage_b200 as g

mode = os.environ.get('MODE', 'x3')
n, d, O = int(os.environ.get('N', 851968)), int(os.environ.get('D', 64)), int(os.environ.get('O', 128))
rows = int(os.environ.get('ROWS', 1632803))
tdt = torch.bfloat16 if mode == 'bf16' else torch.float32
gen = torch.Generator().manual_seed(0)
pad = lambda t: g.ops.pad_table(t, tdt)[0][:, :t.shape[1]]
table = pad(torch.randn((rows, d), generator=gen))
m = pad(torch.randn((n, d), generator=gen))
wx, wn = pad(torch.randn((O, d), generator=gen) / 8), pad(torch.randn((O, d), generator=gen) / 8)
ids = torch.randint(0, rows, (n,), generator=gen).cuda()
out = torch.empty((n, 2 * O), dtype=tdt, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
exact = {'x3': 'x3', 'tf32': False, 'bf16': False}[mode]


def run():
    g.ops.linear([dict(a=table, ids=ids, w=wx, col0=0), dict(a=m, w=wn, col0=O)], n, act='relu', out=out, exact=exact)


for _ in range(3):
    run()
torch.cuda.synchronize()
tot, reps = 0.0, 10
for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
t = tot / reps
es = table.element_size()
byt = n * (2 * table.stride(0) * es + 8 + 2 * O * es)
print('linear %s: n=%d d=%d O=2x%d: %.1f us = %.0f GB/s algorithmic, %.0f TFLOP/s' % (mode, n, d, O, t * 1e3, byt / t / 1e6, 4.0 * n * d * O / t / 1e9))
