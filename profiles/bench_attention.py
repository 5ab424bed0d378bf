#!/usr/bin/env python
"""Micro-benchmark of the attention aggregator's fused reduction (attention_umma.cu) on the plaw2m layer-1 shape: n parents x S = 10
neighbour rows of a 2 M x 256 bf16 table.  CUDA-event timed, L2 flushed.

    python profiles/bench_attention.py        (N / S / D / ROWS change the shape; GSAGE_ATT_TMA_ROWS = rows of a 128-row tile fetched
                                               by TMA gather4 -- the rest come by cp.async; 128 = round 1's all-TMA kernel)"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_graphsage_b200 as g

n, S, d = int(os.environ.get('N', 409600)), int(os.environ.get('S', 10)), int(os.environ.get('D', 256))
rows = int(os.environ.get('ROWS', 2000000))
gen = torch.Generator().manual_seed(0)
table = g.ops.pad_table(torch.randn((rows, d), generator=gen), torch.bfloat16)[0][:, :d]
w1 = g.ops.pad_table(torch.randn((32, d), generator=gen) / 16, torch.bfloat16)[0][:, :d]
w2 = (torch.randn((32, 32), generator=gen) / 6).cuda()
xa = torch.randn((n, 32), generator=gen).cuda()
ids = torch.randint(0, rows, (n * S,), generator=gen).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def run():
    return g.ops.attention_aggregate(table, ids, n, S, w1, w2, xa)


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
ldb = (d + 7) // 8 * 16
byt = n * (S * ldb + 8 * S + ldb + 128)
print('attention_aggregate TMA rows %s: n=%d S=%d d=%d: %.1f us = %.0f GB/s algorithmic  (checksum %.4f)' %
      (os.environ.get('GSAGE_ATT_TMA_ROWS', 'default'), n, S, d, t * 1e3, byt / t / 1e6, out.float().abs().mean().item()))
