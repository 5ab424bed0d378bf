#!/usr/bin/env python
"""Micro-benchmark of the tensor-core projection kernels alone (the layer-1 shape of the reddit workload):
[fc_x(table[ids]) | fc_neib(M)] with n rows, d = 602, O = 128, bf16.  CUDA-event timed.

    python profiles/bench_umma.py [mode ...]     modes: gi (gather + in place, the engine's call; default), ii, gg, g, i
    GSAGE_NO_WS=1 -> streaming kernel (linear_umma.cu); default -> weight-stationary kernel (linear_ws_umma.cu)
    N / D / DT (bf16 | f32) / EXACT (0 = bf16 or single-pass TF32, 1 = FFMA, x3 = 3 x TF32) pick another shape, e.g. the Pokec
    layer-1 call: N=204800 D=64 DT=f32 EXACT=x3"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_graphsage_b200 as g

n, d, O, rows = int(os.environ.get('N', 409600)), int(os.environ.get('D', 602)), 128, 232966
DT = torch.float32 if os.environ.get('DT') == 'f32' else torch.bfloat16
EXACT = {'0': False, '1': True, 'x3': 'x3'}[os.environ.get('EXACT', '0')]
es = 4 if DT == torch.float32 else 2
gen = torch.Generator().manual_seed(0)
table = g.ops.pad_table(torch.randn((rows, d), generator=gen), DT)[0][:, :d]
m = g.ops.pad_table(torch.randn((n, d), generator=gen), DT)[0][:, :d]
m2 = g.ops.pad_table(torch.randn((n, d), generator=gen), DT)[0][:, :d]
wx = g.ops.pad_table(torch.randn((O, d), generator=gen) / 25, DT)[0][:, :d]
wn = g.ops.pad_table(torch.randn((O, d), generator=gen) / 25, DT)[0][:, :d]
ids = torch.randint(0, rows, (n,), generator=gen).cuda()
ids2 = torch.randint(0, rows, (n,), generator=gen).cuda()
out = torch.empty((n, 2 * O), dtype=DT, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
G, I, G2, I2 = dict(a=table, ids=ids, w=wx, col0=0), dict(a=m, w=wn, col0=O), dict(a=table, ids=ids2, w=wn, col0=O), dict(a=m2, w=wx, col0=0)
MODES = {'gi': [G, I], 'ii': [I2, I], 'gg': [G, G2], 'g': [G], 'i': [I]}
for mode in (sys.argv[1:] or ['gi']):
    segs = MODES[mode]
    for _ in range(3):
        g.ops.linear(segs, n, act='relu', out=out, exact=EXACT)
    torch.cuda.synchronize()
    reps, tot = 10, 0.0
    for _ in range(reps):
        flush.zero_()                                   # inputs leave L2 between repetitions
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.ops.linear(segs, n, act='relu', out=out, exact=EXACT)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / reps
    byt = n * len(segs) * ((d + 7) // 8 * 8 * es + O * es) + n * 8 * sum(1 for s in segs if 'ids' in s)
    print('mode=%-3s d=%d %s exact=%s ws=%s stages=%s  %.1f us  %.0f GB/s algorithmic  %.0f TFLOP/s' %
          (mode, d, os.environ.get('DT', 'bf16'), EXACT, '0' if os.environ.get('GSAGE_NO_WS') else '1', os.environ.get('GSAGE_WS_STAGES', '-'), ms * 1e3, byt / ms / 1e6,
           2.0 * len(segs) * n * d * O / ms / 1e9))
