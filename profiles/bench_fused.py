#!/usr/bin/env python
"""Micro-benchmark for the fused kernel (gather_mean_project_umma.cu) against the two launches it replaces, on the
layer-1 shape of the reddit workload: n parents x S = 10 neighbours, d = 602, O = 128, bf16.  CUDA-event timed, L2 flushed.

    python profiles/bench_fused.py            (N=409600 by default; set N / S / D to change the shape)

  unfused:  gather_reduce (mean rows -> HBM)  +  linear [fc_x(table[ids]) | fc_neib(M)]          (what the engine runs today)
  fused:    gather_mean_project (neighbour half, M never leaves the SM)  +  linear [fc_x(table[ids])]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_graphsage_b200 as g

n, S, d, O, rows = int(os.environ.get('N', 409600)), int(os.environ.get('S', 10)), int(os.environ.get('D', 602)), 128, 232966
gen = torch.Generator().manual_seed(0)
table = g.ops.pad_table(torch.randn((rows, d), generator=gen), torch.bfloat16)[0][:, :d]
wx = g.ops.pad_table(torch.randn((O, d), generator=gen) / 25, torch.bfloat16)[0][:, :d]
wn = g.ops.pad_table(torch.randn((O, d), generator=gen) / 25, torch.bfloat16)[0][:, :d]
ids_self = torch.randint(0, rows, (n,), generator=gen).cuda()
ids_nb = torch.randint(0, rows, (n * S,), generator=gen).cuda()
out_a = torch.empty((n, 2 * O), dtype=torch.bfloat16, device='cuda')
out_b = torch.empty((n, 2 * O), dtype=torch.bfloat16, device='cuda')
m = g.ops.pad_table(torch.zeros((n, d)), torch.bfloat16)[0][:, :d]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def unfused():
    g.ops.gather_reduce(table, ids_nb, n, S, 'mean', d=d, out=m)
    g.ops.linear([dict(a=table, ids=ids_self, w=wx, col0=0), dict(a=m, w=wn, col0=O)], n, act='relu', out=out_a, exact=False)


def fused():
    g.ops.gather_mean_project(table, ids_nb, n, S, wn, act='relu', out=out_b, col0=O)
    g.ops.linear([dict(a=table, ids=ids_self, w=wx, col0=0)], n, act='relu', out=out_b, exact=False)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


t_u = timed(unfused)
print('unfused: %.1f us' % (t_u * 1e3))
t_f = timed(fused)
diff = (out_a.float() - out_b.float()).abs().max().item()
ldb = (d + 15) // 16 * 16 * 2
byt = n * (S * ldb + 8 * S + ldb + 8 + 2 * O * 2)
print('fused:   %.1f us  (%.0f GB/s algorithmic)   max |fused - unfused| = %.3g' % (t_f * 1e3, byt / t_f / 1e6, diff))
t_n = timed(lambda: g.ops.gather_mean_project(table, ids_nb, n, S, wn, act='relu', out=out_b, col0=O))
t_g = timed(lambda: g.ops.gather_reduce(table, ids_nb, n, S, 'mean', d=d, out=m))
print('neighbour half alone: fused kernel %.1f us (%.0f GB/s on S rows + ids + out)   gather_reduce %.1f us (%.0f GB/s)' %
      (t_n * 1e3, n * (S * ldb + 8 * S + O * 2) / t_n / 1e6, t_g * 1e3, n * (S * ldb + 8 * S + ldb) / t_g / 1e6))
