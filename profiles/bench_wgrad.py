#!/usr/bin/env python
"""Micro-benchmark of the tensor-core weight-gradient kernel (wgrad_umma.cu) at the layer-1 shape of the reddit workload:
dW (128 x 602) = G^T . A[ids] over n = 26 * 16384 rows.  modes: g (rows by id), i (in place)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_graphsage_b200 as g

n, d, O, rows = int(os.environ.get('N', 425984)), 602, 128, 232966
gen = torch.Generator().manual_seed(0)
table = g.ops.pad_table(torch.randn((rows, d), generator=gen), torch.bfloat16)[0][:, :d]
m = g.ops.pad_table(torch.randn((n, d), generator=gen), torch.bfloat16)[0][:, :d]
G = (torch.randn((n, 2 * O), generator=gen) / 8).to(torch.bfloat16).cuda()
ids = torch.randint(0, rows, (n,), generator=gen).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for mode in (sys.argv[1:] or ['g', 'i']):
    a, i = (table, ids) if mode == 'g' else (m, None)
    for _ in range(2):
        g.ops.wgrad(G[:, :O], a, ids=i, n=n, exact=False)
    torch.cuda.synchronize()
    reps, tot = 10, 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.ops.wgrad(G[:, :O], a, ids=i, n=n, exact=False)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / reps
    byt = n * (608 * 2 + 2 * O * 2)             # A rows once, the G slice once per N-tile (two N-tiles at d = 602)
    print('mode=%s  %.1f us  %.0f GB/s algorithmic  %.0f TFLOP/s' % (mode, ms * 1e3, byt / ms / 1e6, 2.0 * n * d * O / ms / 1e9))
