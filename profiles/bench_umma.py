#!/usr/bin/env python
"""Micro-benchmark of the tensor-core projection kernel alone (the layer-1 shape of the reddit workload):
[fc_x(table[ids]) | fc_neib(M)] with n = 204800 rows, d = 602, O = 128, bf16.  CUDA-event timed."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_graphsage_b200 as g

n, d, O, rows = 204800, 602, 128, 232966
gen = torch.Generator().manual_seed(0)
table = g.ops.pad_table(torch.randn((rows, d), generator=gen), torch.bfloat16)[0][:, :d]
m = g.ops.pad_table(torch.randn((n, d), generator=gen), torch.bfloat16)[0][:, :d]
wx = g.ops.pad_table(torch.randn((O, d), generator=gen) / 25, torch.bfloat16)[0][:, :d]
wn = g.ops.pad_table(torch.randn((O, d), generator=gen) / 25, torch.bfloat16)[0][:, :d]
ids = torch.randint(0, rows, (n,), generator=gen).cuda()
out = torch.empty((n, 2 * O), dtype=torch.bfloat16, device='cuda')
segs = [dict(a=table, ids=ids, w=wx, col0=0), dict(a=m, w=wn, col0=O)]
for _ in range(3):
    g.ops.linear(segs, n, act='relu', out=out, exact=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    g.ops.linear(segs, n, act='relu', out=out, exact=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
byt = n * (2 * 608 * 2 + 8 + 2 * O * 2)
print('debug=%s  %.1f us  %.0f GB/s algorithmic  %.0f TFLOP/s' % (os.environ.get('GSAGE_UMMA_DEBUG', '0'), ms * 1e3, byt / ms / 1e6,
                                                                 4.0 * n * d * O / ms / 1e9))
