#!/usr/bin/env python
"""
bench.py -- the headline measurement (BASELINE.json metric: sampled+aggregated nodes/sec; gather HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload reddit] [--batch B]

A "step" is one pass of the hot path (sample both hops -> gather -> aggregate -> project, two layers, normalise,
classifier) over one batch of B seed nodes per GPU on a synthetic problem of the named shape.  Default workload =
BASELINE.json configs[1]: Reddit-shape (232,965 nodes / 11 M edges / d=602), mean aggregator, bf16 table,
fanout [25,10], on 1 x B200.

Ours (`--impl ours`):
  value   sampled neighbour rows consumed by aggregation per second (= seeds/s x 275), whole job over N GPUs,
          ids already resident in HBM, device-timed (CUDA events, max over ranks)
  e2e     same metric through the host-buffer entry (pinned host ids in, pinned host logits out every step)
  roofline  the fused gather+aggregate kernel: algorithmic bytes / live CUDA-event time vs measured HBM peak
  cpu_baseline  the oracle port of the reference's CPU path, bounded sample, timed here on the host cores
Reference arm (`--impl reference`): the oracle port of the reference's CPU implementation (the reference is
Python-2-era and cannot travel to the GPU box; oracle/ restates it and is pinned against it by the golden
fixtures) timed on the host cores with all torch threads.
"""

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RANK = int(os.environ.get('RANK', '0'))
LOCAL_RANK = int(os.environ.get('LOCAL_RANK', '0'))
WORLD = int(os.environ.get('WORLD_SIZE', '1'))

FANOUT = (25, 10)
OUT_DIMS = (128, 128)
ROWS_PER_SEED = FANOUT[0] + FANOUT[0] * FANOUT[1]          # 275 sampled neighbour rows per seed
METRIC = 'sampled+aggregated nodes/sec'
UNIT = 'nodes/s'

WORKLOADS = {
    # name -> (synth shape, aggregator, prep, table dtype, with_feats)
    'reddit': ('reddit', 'mean', 'identity', 'bf16', True),          # BASELINE.json configs[1]  (default)
    'reddit-fp32': ('reddit', 'mean', 'identity', 'f32', True),
    'pokec-mean': ('pokec', 'mean', 'node_embedding', 'f32', False),  # north-star 60 % target shape
    'pokec-mean-tf32': ('pokec', 'mean', 'node_embedding', 'tf32', False),    # same, projections as TF32 on tcgen05
    'pokec-maxpool': ('pokec', 'max_pool', 'node_embedding', 'bf16', False),  # configs[2] (bf16 compute: the MLP is tensor-bound)
    'reddit-maxpool': ('reddit', 'max_pool', 'identity', 'bf16', True),      # max-pool on the reddit shape (trainable: identity prep)
    'plaw2m-attention': ('plaw2m', 'attention', 'identity', 'bf16', True),   # configs[3]
    'big10m': ('big10m', 'mean', 'identity', 'bf16', True),                   # configs[4]
    'reddit-lstm': ('reddit', 'lstm', 'identity', 'bf16', True),              # LSTM aggregator (forward only; use --batch 2048)
    'tiny': ('tiny', 'mean', 'identity', 'f32', True),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='reddit', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=16384, help='seed nodes per step per GPU')
    ap.add_argument('--cpu-batch', type=int, default=512, help='seed nodes per CPU-baseline step (train.py:48)')
    ap.add_argument('--cpu-seconds', type=float, default=12.0, help='budget of the cpu_baseline leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true', help='skip the train-step leg')
    ap.add_argument('--torch-adam', action='store_true', help='train leg: torch.optim.Adam + clip_grad_norm_ instead of the fused native step')
    ap.add_argument('--no-ahead', action='store_true', help='sample inside each forward instead of one batch ahead on the sampler stream')
    ap.add_argument('--scale', type=float, default=1.0, help='shrink the graph (debug only; reported in config)')
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, '/tmp/gsage_clocks_%d.csv' % os.getpid()

    def start(self):
        try:
            self.fh = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '20'], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons, power = [], [], set(), []
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        os.remove(self.path)
        sm.sort()
        return {'sm_mhz': (sm[len(sm) // 2] if sm else None), 'sm_max_mhz': (max(mx) if mx else None),
                'power_w_max': (max(power) if power else None), 'samples': len(sm), 'reasons': sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
def make_problem(args):
    from pytorch_graphsage_b200 import synth
    shape, agg, prep, dtype, with_feats = WORKLOADS[args.workload]
    t0 = time.time()
    prob = synth.make_problem(shape, seed=0, with_feats=with_feats, scale=args.scale)
    prob.update(aggregator=agg, prep=prep, table_dtype=dtype, build_s=time.time() - t0)
    return prob


def layer_specs():
    from torch.nn import functional as F
    return [dict(n_train_samples=FANOUT[0], n_val_samples=FANOUT[0], output_dim=OUT_DIMS[0], activation=F.relu),
            dict(n_train_samples=FANOUT[1], n_val_samples=FANOUT[1], output_dim=OUT_DIMS[1], activation=lambda x: x)]


def workload_config(args, prob):
    s = prob['adj']
    return {'workload': '%s: %d nodes / %d edges / d=%s, %s aggregator, %s prep, fanout [25,10], out 128,128, %s table' %
                        (args.workload, s['n_nodes'], s['nnz'], prob['feats_dim'] or 64, prob['aggregator'], prob['prep'],
                         prob['table_dtype']),
            'batch_seeds_per_gpu': args.batch, 'rows_per_seed': ROWS_PER_SEED, 'graph_scale': args.scale,
            'sampler': 'sparse_uniform_neighbor_sampler (device MT19937, bit-exact numpy legacy stream)',
            'l2_policy': 'inputs larger than L2 (table %.0f MB + per-step gather footprint); no flush' %
                         ((s['n_nodes'] + 1) * (prob['feats_dim'] or 64) * (2 if prob['table_dtype'] == 'bf16' else 4) / 1e6),
            'pipeline': ('none: every forward samples its own batch' if getattr(args, 'no_ahead', False) else
                         'sample-ahead: the draws + CSR lookups of batch i+1 run on a second stream under the aggregation of batch i; '
                         'every timed step still samples one batch and aggregates one batch'),
            'parallelism': 'seed-sharded dp%d, graph+table replicated, no data-path collective' % max(1, args.gpus)}


# ---------------------------------------------------------------------------------------------------
def cpu_reference_throughput(prob, batch, steps, warmup, seconds=None):
    """The reference's CPU path (oracle port: numpy legacy RNG + CSR lookup + torch-CPU fp32 layers) on `batch`
    seeds per step.  Returns (rows_per_s, ms_per_step, steps_done).  bf16 workloads use the bf16-rounded table
    upcast to fp32 (the reference has no bf16 path)."""
    import numpy as np
    import torch
    from oracle import layers, sampler as osampler
    from pytorch_graphsage_b200 import synth
    host_threads()
    adj = prob['adj']
    indptr, data, shape = adj['indptr'], adj['data'], adj['shape']
    indices = np.arange(data.shape[0], dtype=np.int64) - np.repeat(indptr[:-1], np.diff(indptr))
    deg = np.diff(indptr)
    feats = None
    if prob['feats'] is not None:
        feats = torch.from_numpy(prob['feats'])
        if prob['table_dtype'] == 'bf16':
            feats = feats.to(torch.bfloat16).float()
    params = reference_params(prob)
    rs = np.random.RandomState(123 ** 2)
    draw = lambda hi, n: rs.choice(hi, n)
    done, t_start, times = 0, time.perf_counter(), []
    with torch.no_grad():
        while True:
            ids0 = synth.seed_batch(prob, batch, seed=1000 + done)
            t0 = time.perf_counter()
            ids1 = osampler.sparse_sample(indptr, indices, data, shape, deg, ids0, FANOUT[0], draw)
            ids2 = osampler.sparse_sample(indptr, indices, data, shape, deg, ids1, FANOUT[1], draw)
            layers.forward_stack([torch.from_numpy(a) for a in (ids0, ids1, ids2)], feats, params,
                                 aggregator=prob['aggregator'], prep=prob['prep'], n_nodes=prob['n_nodes'])
            dt = time.perf_counter() - t0
            done += 1
            if done > warmup:
                times.append(dt)
            if seconds is not None:
                if len(times) >= 3 and time.perf_counter() - t_start > seconds:
                    break
            elif len(times) >= steps:
                break
    ms = 1e3 * sum(times) / len(times)
    return batch * ROWS_PER_SEED / (ms / 1e3), ms, len(times)


def host_threads():
    """All the host threads the CPU path can use.  torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs
    run on rank 0 alone, so give them the box (physical cores = half the logical CPUs, torch's own default)."""
    import torch
    want = max(1, (os.cpu_count() or 2) // 2)
    if torch.get_num_threads() < want:
        torch.set_num_threads(want)
    return torch.get_num_threads()


def reference_params(prob, seed=123):
    """Random-init weights of the reference architecture (torch.manual_seed(123), train.py:81), keyed like its state_dict."""
    import torch
    from torch import nn
    torch.manual_seed(seed)
    d = prob['feats_dim']
    params = {}
    if prob['prep'] == 'node_embedding':
        params['prep.embedding.weight'] = nn.Embedding(prob['n_nodes'] + 1, 64).weight.data
        fc = nn.Linear(64, 64)
        params['prep.fc.weight'], params['prep.fc.bias'] = fc.weight.data, fc.bias.data
        d = (d or 0) + 64
    elif prob['prep'] == 'linear':
        params['prep.fc.weight'] = nn.Linear(d, 32, bias=False).weight.data
        d = 32
    for k, O in enumerate(OUT_DIMS):
        pre = 'agg_layers.%d.' % k
        hid = d
        if prob['aggregator'] in ('max_pool', 'mean_pool'):
            mlp = nn.Linear(d, 512)
            params[pre + 'mlp.0.weight'], params[pre + 'mlp.0.bias'] = mlp.weight.data, mlp.bias.data
            hid = 512
        if prob['aggregator'] == 'lstm':
            lstm = nn.LSTM(d, 512, batch_first=True)
            for name in ('weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0'):
                params[pre + 'lstm.' + name] = getattr(lstm, name).data
            hid = 512
        if prob['aggregator'] == 'attention':
            params[pre + 'att.0.weight'] = nn.Linear(d, 32, bias=False).weight.data
            params[pre + 'att.2.weight'] = nn.Linear(32, 32, bias=False).weight.data
        params[pre + 'fc_x.weight'] = nn.Linear(d, O, bias=False).weight.data
        params[pre + 'fc_neib.weight'] = nn.Linear(hid, O, bias=False).weight.data
        d = 2 * O
    fc = nn.Linear(d, prob['n_classes'])
    params['fc.weight'], params['fc.bias'] = fc.weight.data, fc.bias.data
    return params


def run_reference(args):
    """`--impl reference`: rank 0 alone times the CPU path; other ranks exit 0 without work."""
    if RANK != 0:
        return
    import torch
    prob = make_problem(args)
    value, ms, steps = cpu_reference_throughput(prob, args.cpu_batch, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': dict(workload_config(args, prob), batch_seeds_per_step=args.cpu_batch, pipeline='n/a (the reference\'s CPU path)'),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': '%d steps x %d seeds (oracle port of nn_modules.py/models.py CPU path, torch %d threads of %d cpus)' %
                                       (steps, args.cpu_batch, cores, os.cpu_count())},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'seeds_per_s': value / ROWS_PER_SEED, 'gpu_launches': 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import pytorch_graphsage_b200 as g
    from pytorch_graphsage_b200 import synth

    torch.cuda.set_device(LOCAL_RANK)                                         # one process per GPU
    if WORLD > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() not in ('INFO', 'TRACE'):
            os.environ['NCCL_DEBUG'] = 'WARN'                                  # keep stdout to the one JSON line
        dist.init_process_group('nccl', device_id=torch.device('cuda', LOCAL_RANK))

    prob = make_problem(args)
    B = args.batch
    dtype = torch.bfloat16 if prob['table_dtype'] == 'bf16' else torch.float32
    tf32 = prob['table_dtype'] == 'tf32'
    graph = g.GraphCSR.from_synth(prob['adj'])
    table = g.FeatureTable(prob['feats'], dtype) if prob['feats'] is not None else None
    model = g.GSSupervised(
        input_dim=prob['feats_dim'], n_nodes=prob['n_nodes'], n_classes=prob['n_classes'], layer_specs=layer_specs(),
        aggregator_class=g.aggregator_lookup[prob['aggregator']], prep_class=g.prep_lookup[prob['prep']],
        sampler_class=g.sampler_lookup['sparse_uniform_neighbor_sampler'], adj=graph, train_adj=graph,
        compute_dtype=dtype, max_batch=B, allow_tf32=tf32)
    model.load_state_dict(reference_params(prob))
    model = model.cuda()
    g.set_seeds(123 ** 2 + RANK)                                              # train.py:133 (+ rank: disjoint streams)

    n_batches = 8                                                               # rotate seed batches: no step reuses hot rows
    dev_ids = [torch.from_numpy(synth.seed_batch(prob, B, seed=17 * RANK + i)).cuda() for i in range(n_batches)]
    host_ids = [torch.from_numpy(synth.seed_batch(prob, B, seed=17 * RANK + i)).pin_memory() for i in range(n_batches)]
    host_out = torch.empty((B, prob['n_classes']), dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if WORLD > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if WORLD == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing -----------------------------------------------------------------------------
    ahead = not args.no_ahead
    step_no = [0]

    def step():
        i = step_no[0]
        step_no[0] += 1
        out = model(dev_ids[i % n_batches], table)
        if ahead:                                                                # batch i+1 is drawn while batch i aggregates
            model.sample_ahead(dev_ids[(i + 1) % n_batches], table)
        return out

    if ahead:
        model.sample_ahead(dev_ids[0], table)
    for i in range(max(3, args.warmup)):
        step()
    g.default_rng().check()
    graph.check()
    model.profile(True)
    try:
        gpu_sel = 'GPU-' + str(torch.cuda.get_device_properties(LOCAL_RANK).uuid)
    except Exception:
        gpu_sel = str(LOCAL_RANK)
    clocks = ClockSampler(gpu_sel)
    barrier()
    clocks.start()
    launches0 = g.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step()
    ev1.record()
    barrier()
    launches = g.launch_count() - launches0
    clk = clocks.stop()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    prof = model.profile_read()
    g.default_rng().check()
    ms_step = ms_total / args.steps
    value = WORLD * B * ROWS_PER_SEED / (ms_step / 1e3)

    # the dominant kernel alone: a few steps WITHOUT sample-ahead, so that no sampler kernel shares the SMs with it
    # (in the pipelined run above its launches overlap the next batch's draws: the step gets shorter, the kernel longer)
    prof_iso = None
    if ahead:
        model(dev_ids[step_no[0] % n_batches], table)                          # consume the batch the last timed step drew ahead
        step_no[0] += 1
        torch.cuda.synchronize()
        model.profile_read()
        for i in range(20):
            model(dev_ids[(step_no[0] + i) % n_batches], table)
        torch.cuda.synchronize()
        prof_iso = model.profile_read()
        step_no[0] += 20
    model.profile(False)

    # ---- end to end through the host-buffer entry -----------------------------------------------------------------
    def host_step(i):                                                          # synchronises every step (D2H of the logits)
        nxt = host_ids[(i + 1) % n_batches] if ahead else None
        model.forward_host(host_ids[i % n_batches], table, host_out, next_ids_host=nxt)

    for i in range(3):
        host_step(i)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for i in range(3, 3 + args.steps):
        host_step(i)
    ev1.record()
    barrier()
    if ahead:
        model.forward_host(host_ids[(3 + args.steps) % n_batches], table, host_out)   # drain the pending batch
    e2e_ms = max_over_ranks(max(ev0.elapsed_time(ev1), 0.0)) / args.steps
    e2e_wall_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    e2e_ms = max(e2e_ms, e2e_wall_ms)
    e2e_value = WORLD * B * ROWS_PER_SEED / (e2e_ms / 1e3)

    # ---- one optimiser step per batch (forward + loss + backward + gradient all-reduce + clip + Adam) ---------------
    train = None
    trainable = (prob['aggregator'] == 'mean' and (prob['prep'] == 'identity' or (prob['prep'] == 'node_embedding' and dtype == torch.float32 and not tf32))) or \
                (prob['aggregator'] in ('max_pool', 'mean_pool') and dtype == torch.bfloat16) or \
                (prob['aggregator'] == 'attention' and prob['prep'] == 'identity' and dtype == torch.bfloat16)
    if trainable and not args.no_train:
        from torch.nn import functional as F
        if prob['task'] == 'regression_mae':                                       # problem.py:39-41: l1 loss on (B, 1) predictions
            tgt_all = torch.from_numpy(prob['targets']).cuda()
            loss_fn = lambda preds, t: F.l1_loss(preds, t.view_as(preds))
        else:
            tgt_all = torch.from_numpy(prob['targets'].reshape(-1)).cuda()
            loss_fn = F.cross_entropy
        tgts = [tgt_all[i] for i in dev_ids]
        # clip_grad_norm 5 + Adam as one native call over flat buffers (parallel.FusedAdam); --torch-adam: the stock pair
        opt = torch.optim.Adam(model.parameters(), lr=0.01) if args.torch_adam else g.FusedAdam(model, lr=0.01)
        side = torch.cuda.Stream()
        k_train = max(3, min(args.steps, 30))
        def train_step(i):
            model.train_step(dev_ids[i % n_batches], table, tgts[i % n_batches], loss_fn, optimizer=opt, grad_scale=1.0 / WORLD,
                             overlap_stream=side, next_ids=dev_ids[(i + 1) % n_batches] if ahead else None)

        for i in range(3):
            train_step(i)
        barrier()
        ev0.record()
        for i in range(3, 3 + k_train):
            train_step(i)
        ev1.record()
        barrier()
        if ahead:
            model(dev_ids[(3 + k_train) % n_batches], table)                       # drain the pending batch
        t_ms = max_over_ranks(ev0.elapsed_time(ev1)) / k_train
        train = {'ms_per_step': t_ms, 'seeds_per_s': WORLD * B / (t_ms / 1e3), 'steps': k_train,
                 'allreduce_bytes_per_step': int(model._bucket().flat.numel()) * 4,
                 'collective': 'one flat fp32 gradient bucket, NCCL all-reduce in two pieces (fc + layer-2 head overlapped with the '
                               'layer-1 weight-gradient kernels)' if WORLD > 1 else 'none (1 GPU)',
                 'note': 'loss is stock torch; clip + Adam are ' + ('stock torch' if args.torch_adam else 'one native call (gsage_adam_step)')}

    if RANK != 0:
        if WORLD > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    red_ms, red_n, red_bytes = prof['reduce']
    prj_ms, prj_n, prj_flops = prof['project']
    achieved = (red_bytes / 1e9) / (red_ms / 1e3) if red_ms > 0 else None
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get('%s:%d' % (args.workload, B))
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': WORLD, 'steps': args.steps, 'warmup': max(3, args.warmup),
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if dtype == torch.bfloat16 else ('tf32' if tf32 else 'f32'), 'data': 'synthetic',
        'config': workload_config(args, prob), 'clocks': clk,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': 8 * B, 'd2h_bytes_per_step': 4 * B * prob['n_classes'],
                'ms_per_step': e2e_ms},
        'gpu_launches': int(launches) * WORLD,
        'seeds_per_s': value / ROWS_PER_SEED,
        'roofline': {'bound': 'hbm', 'kernel': 'gather_reduce_kernel, the fused gather+mean launch of layer 1 on the (x1, x2) pair (B*25 parents, S=10)',
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': (achieved / peak if achieved else None),
                     'traffic': traffic, 'peak_source': peak_src, 'launches': int(red_n),
                     'algorithmic_bytes_per_launch_avg': (red_bytes / red_n if red_n else None),
                     'avg_launch_ms': (red_ms / red_n if red_n else None),
                     'timed_in': 'the pipelined steps above (its launches share the SMs with the next batch\'s sampling kernels)' if ahead
                                 else 'the timed steps above'},
        'breakdown_ms_per_step': {'forward': prof['forward'][0] / args.steps, 'sample': prof['sample'][0] / args.steps,
                                  'gather_reduce': red_ms / args.steps, 'project': prj_ms / args.steps,
                                  # fused build: the projection runs inside the gather+aggregate kernel (no time of its own)
                                  'project_tflops': (prj_flops / 1e12) / ((prj_ms if prj_ms > 0 else red_ms) / 1e3) if (prj_ms + red_ms) > 0 else None},
    }
    if prof_iso is not None and prof_iso['reduce'][0] > 0:
        ims, inn, iby = prof_iso['reduce']
        line['roofline']['isolated'] = {'achieved': (iby / 1e9) / (ims / 1e3), 'frac': (iby / 1e9) / (ims / 1e3) / peak, 'launches': int(inn),
                                        'avg_launch_ms': ims / inn,
                                        'note': 'same kernel, same inputs, 20 steps without sample-ahead (nothing else on the SMs)'}
    if train is not None:
        line['train'] = train
    if not args.no_cpu_baseline:
        v, ms, steps = cpu_reference_throughput(prob, args.cpu_batch, None, 2, seconds=args.cpu_seconds)
        line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                                'sample': '%d steps x %d seeds in %.0f s budget (oracle port, torch %d threads of %d cpus)' %
                                          (steps, args.cpu_batch, args.cpu_seconds, torch.get_num_threads(), os.cpu_count()),
                                'ms_per_step': ms}
    print(json.dumps(line))
    if WORLD > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
