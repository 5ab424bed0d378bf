#!/usr/bin/env python
"""A few optimiser steps of the bench workload and nothing else -- the command profiled for the train-step launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv \
        python profiles/train_probe.py [steps] [batch]"""
import os
import sys
import torch
from torch.nn import functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                     # noqa: E402  (workload definition + reference-initialised weights)
import pytorch_graphsage_b200 as g               # noqa: E402
from pytorch_graphsage_b200 import synth         # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
workload = sys.argv[3] if len(sys.argv) > 3 else 'reddit'
args = type('A', (), dict(workload=workload, scale=1.0))()
prob = bench.make_problem(args)
graph = g.GraphCSR.from_synth(prob['adj'])
table = g.FeatureTable(prob['feats'], torch.bfloat16)
model = g.GSSupervised(input_dim=prob['feats_dim'], n_nodes=prob['n_nodes'], n_classes=prob['n_classes'], layer_specs=bench.layer_specs(),
                       aggregator_class=g.aggregator_lookup[prob['aggregator']], prep_class=g.prep_lookup[prob['prep']],
                       sampler_class=g.sampler_lookup['sparse_uniform_neighbor_sampler'], adj=graph, train_adj=graph,
                       compute_dtype=torch.bfloat16, max_batch=B)
model.load_state_dict(bench.reference_params(prob))
model = model.cuda()
g.set_seeds(123 ** 2)
ids = [torch.from_numpy(synth.seed_batch(prob, B, seed=i)).cuda() for i in range(4)]
tgt = torch.from_numpy(prob['targets'].reshape(-1)).cuda()
opt = g.FusedAdam(model, lr=0.01)
side = torch.cuda.Stream()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(steps + 2):
    if i == 2:
        torch.cuda.synchronize(); ev0.record()
    model.train_step(ids[i % 4], table, tgt[ids[i % 4]], F.cross_entropy, optimizer=opt, overlap_stream=side,
                     next_ids=ids[(i + 1) % 4])
ev1.record()
torch.cuda.synchronize()
print('train step: %.3f ms' % (ev0.elapsed_time(ev1) / steps))
