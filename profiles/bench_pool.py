#!/usr/bin/env python
"""Micro-benchmark of the pool aggregators' fused MLP + pool launch (linear_pool_ws_umma.cu) on the pokec layer-1 shape:
n parents x S = 10 neighbour rows of a 1.63 M x 64 bf16 table, 512 hidden units.  CUDA-event timed, L2 flushed.

    python profiles/bench_pool.py                 (N / S / D / H / ROWS change the shape; GSAGE_POOL_PIPE=0: unpipelined epilogue;
                                                   GSAGE_B200_LIB=.../libgsage_b200_timing.so: per-role cycle counters)"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_graphsage_b200 as g

n, S, d, H = int(os.environ.get('N', 409600)), int(os.environ.get('S', 10)), int(os.environ.get('D', 64)), int(os.environ.get('H', 512))
rows = int(os.environ.get('ROWS', 1632803))
gen = torch.Generator().manual_seed(0)
table = g.ops.pad_table(torch.randn((rows, d), generator=gen), torch.bfloat16)[0][:, :d]
w = g.ops.pad_table(torch.randn((H, d), generator=gen) / 8, torch.bfloat16)[0][:, :d]
bias = (torch.randn((H,), generator=gen) / 8).cuda()
ids = torch.randint(0, rows, (n * S,), generator=gen).cuda()
seq = os.environ.get('SEQ') == '1'                      # rows read in place (no ids): isolates the gather from everything else
dense = table[:n * S] if seq and rows >= n * S else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def run():
    if dense is not None:
        return g.ops.linear_pooled(dense, w, n, S, 'max', bias=bias, out_dtype=torch.bfloat16)
    return g.ops.linear_pooled(table, w, n, S, 'max', ids=ids, bias=bias, out_dtype=torch.bfloat16)


for _ in range(3):
    out = run()
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
flop = 2.0 * n * S * d * H
print('linear_pooled %s: n=%d S=%d d=%d H=%d: %.1f us = %.0f TFLOP/s, %.2f G neighbour rows/s' % ('in place' if dense is not None else 'gathered', n, S, d, H, t * 1e3, flop / t / 1e9, n * S / t / 1e6))
