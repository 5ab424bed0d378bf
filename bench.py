#!/usr/bin/env python
"""
bench.py -- the headline measurement (BASELINE.json metric: sampled+aggregated nodes/sec; gather HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload reddit] [--batch B] [--legs ...]

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
  configs the other BASELINE.json configs as short legs of the same run (`--legs`, default all): pokec-mean (the north-star
          60 % target shape), pokec-maxpool (configs[2], bf16 and tf32), plaw2m-attention (configs[3]), big10m (configs[4]) --
          each with value, ms/step, the dominant kernel's roofline and, where a backward exists, a train-step leg (forward +
          loss + backward + gradient all-reduce over the N ranks + clip + Adam)
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
    'pokec-maxpool-tf32': ('pokec', 'max_pool', 'node_embedding', 'tf32', False),   # configs[2] on fp32 tables (TF32 tensor-core products)
    'pokec-maxpool-f32': ('pokec', 'max_pool', 'node_embedding', 'f32', False),     # configs[2] at the reference's precision (3 x TF32: fp32-exact)
    'reddit-maxpool': ('reddit', 'max_pool', 'identity', 'bf16', True),      # max-pool on the reddit shape (trainable: identity prep)
    'plaw2m-attention': ('plaw2m', 'attention', 'identity', 'bf16', True),   # configs[3]
    'plaw2m-attention-tf32': ('plaw2m', 'attention', 'identity', 'tf32', True),   # configs[3] on an fp32 table (TF32 products)
    'plaw2m-attention-f32': ('plaw2m', 'attention', 'identity', 'f32', True),     # configs[3] at the reference's precision (3 x TF32: fp32-exact)
    'big10m': ('big10m', 'mean', 'identity', 'bf16', True),                   # configs[4]
    'reddit-lstm': ('reddit', 'lstm', 'identity', 'bf16', True),              # LSTM aggregator (forward only; use --batch 2048)
    'tiny': ('tiny', 'mean', 'identity', 'f32', True),
}

# the legs of the default run: (workload, seeds per step per GPU, timed steps); the first is the headline line
DEFAULT_LEGS = [('pokec-mean', 32768, 40), ('pokec-maxpool', 16384, 40), ('plaw2m-attention', 16384, 40), ('big10m', 16384, 40),
                # configs[2] / configs[3] again at the reference's precision: fp32 tables and activations, every product 3 x TF32
                # (~1e-6 relative, the golden fixtures' 1e-4 bar); BASELINE.json states bf16 only for configs[1]
                ('pokec-maxpool-f32', 4096, 10), ('plaw2m-attention-f32', 4096, 10)]
# workloads whose feature table is generated on the device (a host copy would be 10 GB of fp32 for big10m)
DEVICE_FEATS = ('plaw2m', 'big10m')


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
    ap.add_argument('--legs', default='default', help="'default' (the other BASELINE configs as short legs), 'none', or a comma list of "
                                                      "workload[:batch[:steps]]; only with the default workload")
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


def measured_tensor_peak():
    """Dense bf16 TFLOP/s for a kernel timed inside a long step: the SUSTAINED cuBLAS figure of MEASURED_PEAKS.json."""
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            return float(p['bf16_tflops_sustained']), 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)'
        except Exception:
            pass
    return 1500.0, 'fallback (B200_PROFILING.md ~1.5 PFLOP/s sustained dense bf16)'


class ClockSampler(object):
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md clocks line).  NVML from a thread of this process
    (pynvml ships with the image): an `nvidia-smi -lms` child needs a second or more to come up on an 8-GPU box, longer than the
    timed region under torchrun -- round 1's SCALE lines carried 0 samples for that reason.  Falls back to nvidia-smi."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, uuid, index):
        self.uuid, self.index = uuid, index
        self.proc, self.thread, self.stop_flag = None, None, False
        self.sm, self.reasons, self.power, self.sm_max = [], set(), [], None
        self.path = '/tmp/gsage_clocks_%d.csv' % os.getpid()
        self.how = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(self.uuid.encode() if isinstance(self.uuid, str) else self.uuid)
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def prepare(self):
        """Everything slow (NVML init) happens here, before the barrier in front of the timed region."""
        try:
            self.nv, self.h = self._nvml_handle()
            self.sm_max = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.how = 'nvml'
        except Exception:
            self.how = 'nvidia-smi'

    def _loop(self):
        nv, h = self.nv, self.h
        names = (('hw_slowdown', getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8)),
                 ('hw_thermal_slowdown', getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40)),
                 ('sw_thermal_slowdown', getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20)),
                 ('sw_power_cap', getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4)))
        reasons_fn = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons')
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1e3)
                bits = int(reasons_fn(h))
                for name, bit in names:
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.how is None:
            self.prepare()
        if self.how == 'nvml':
            import threading
            self.stop_flag = False
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            self.fh = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.uuid), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '20'], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.how == 'nvml':
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm = sorted(self.sm)
            return {'sm_mhz': (sm[len(sm) // 2] if sm else None), 'sm_max_mhz': self.sm_max,
                    'power_w_max': (max(self.power) if self.power else None), 'samples': len(sm), 'reasons': sorted(self.reasons),
                    'source': 'nvml thread, 4 ms period, inside the timed region'}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable'], 'samples': 0}
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
                'power_w_max': (max(power) if power else None), 'samples': len(sm), 'reasons': sorted(reasons), 'source': 'nvidia-smi -lms 20'}


# ---------------------------------------------------------------------------------------------------
def make_problem(workload, scale=1.0, host_feats=True):
    """The synthetic problem of a workload.  `host_feats` False: the feature table of the big shapes is left out here and drawn
    on the device (device_table) -- same distribution (N(0,1), row 0 = zeros), different generator."""
    from pytorch_graphsage_b200 import synth
    shape, agg, prep, dtype, with_feats = WORKLOADS[workload]
    t0 = time.time()
    on_device = with_feats and not host_feats and shape in DEVICE_FEATS
    prob = synth.make_problem(shape, seed=0, with_feats=with_feats and not on_device, scale=scale)
    if on_device:
        prob['feats_dim'] = synth.SHAPES[shape]['d']
    prob.update(aggregator=agg, prep=prep, table_dtype=dtype, build_s=time.time() - t0, feats_on_device=on_device, workload=workload)
    return prob


def device_table(prob, dtype):
    """(n_nodes, d) N(0,1) feature table drawn on the GPU in row chunks straight into the padded layout (row 0 = the dummy's zeros)."""
    import torch
    import pytorch_graphsage_b200 as g
    rows, d = prob['n_nodes'], prob['feats_dim']
    per = 16 // (2 if dtype == torch.bfloat16 else 4)
    ld = (d + 2 * per - 1) // (2 * per) * (2 * per)
    store = torch.zeros((rows, ld), dtype=dtype, device='cuda')
    gen = torch.Generator(device='cuda').manual_seed(7919)
    step = 1 << 20
    for lo in range(0, rows, step):
        hi = min(rows, lo + step)
        store[lo:hi, :d] = torch.randn((hi - lo, d), generator=gen, device='cuda', dtype=torch.float32).to(dtype)
    store[0].zero_()
    return g.FeatureTable.from_store(store, d)


def layer_specs():
    from torch.nn import functional as F
    return [dict(n_train_samples=FANOUT[0], n_val_samples=FANOUT[0], output_dim=OUT_DIMS[0], activation=F.relu),
            dict(n_train_samples=FANOUT[1], n_val_samples=FANOUT[1], output_dim=OUT_DIMS[1], activation=lambda x: x)]


def workload_config(workload, prob, batch, gpus, scale=1.0, ahead=True):
    s = prob['adj']
    return {'workload': '%s: %d nodes / %d edges / d=%s, %s aggregator, %s prep, fanout [25,10], out 128,128, %s table' %
                        (workload, s['n_nodes'], s['nnz'], prob['feats_dim'] or 64, prob['aggregator'], prob['prep'],
                         prob['table_dtype']),
            'batch_seeds_per_gpu': batch, 'rows_per_seed': ROWS_PER_SEED, 'graph_scale': scale,
            'sampler': 'sparse_uniform_neighbor_sampler (device MT19937, bit-exact numpy legacy stream)',
            'l2_policy': 'inputs larger than L2 (table %.0f MB + per-step gather footprint); no flush' %
                         ((s['n_nodes'] + 1) * (prob['feats_dim'] or 64) * (2 if prob['table_dtype'] == 'bf16' else 4) / 1e6),
            'pipeline': ('none: every forward samples its own batch' if not ahead else
                         'sample-ahead: the draws + CSR lookups of batch i+1 run on a second stream under the aggregation of batch i; '
                         'every timed step still samples one batch and aggregates one batch'),
            'parallelism': 'seed-sharded dp%d, graph+table replicated, no data-path collective' % max(1, gpus)}


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
    prob = make_problem(args.workload, args.scale, host_feats=True)
    value, ms, steps = cpu_reference_throughput(prob, args.cpu_batch, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': dict(workload_config(args.workload, prob, args.batch, args.gpus, args.scale), batch_seeds_per_step=args.cpu_batch,
                           pipeline='n/a (the reference\'s CPU path)'),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': '%d steps x %d seeds (oracle port of nn_modules.py/models.py CPU path, torch %d threads of %d cpus)' %
                                       (steps, args.cpu_batch, cores, os.cpu_count())},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'seeds_per_s': value / ROWS_PER_SEED, 'gpu_launches': 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def barrier():
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if WORLD > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x):
    import torch
    import torch.distributed as dist
    if WORLD == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def dominant_roofline(prob, prof, label_timed_in, traffic=None):
    """The `roofline` object of a leg from the engine's live stopwatch (GSAGE_PROF_REDUCE = the dominant aggregate launch)."""
    ms, n, byt, flo = prof['reduce']
    if not n or ms <= 0:
        return None
    agg = prob['aggregator']
    if agg in ('max_pool', 'mean_pool'):
        peak, src = measured_tensor_peak()
        dtype_note = ''
        kernel = 'linear_pool_ws_umma_kernel: relu(W1 n + b1) on tcgen05 + the pool over the S rows in the epilogue'
        if prob['table_dtype'] != 'bf16':
            peak, dtype_note = peak / 2.0, ' / 2 (TF32 runs at half the bf16 rate)'
        if prob['table_dtype'] == 'f32':
            kernel = ('the MLP as 3 x TF32 projections (linear_ws_umma_kernel in 128-column blocks: three tensor-core products per algorithmic '
                      'one, fp32-exact) + the pool as a segment reduce over the hidden rows')
        kernel += ', layer 1 on the (x1, x2) pair (B*25 parents, S=10)'
        achieved = (flo / 1e12) / (ms / 1e3)
        return {'bound': 'tensor', 'kernel': kernel,
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
                'peak_source': src + dtype_note, 'launches': int(n), 'algorithmic_flops_per_launch_avg': flo / n,
                'algorithmic_bytes_per_launch_avg': byt / n, 'hbm_gbs_algorithmic': (byt / 1e9) / (ms / 1e3), 'avg_launch_ms': ms / n,
                'timed_in': label_timed_in}
    peak, src = measured_peaks()
    achieved = (byt / 1e9) / (ms / 1e3)
    kernel = (('attention_fused_kernel: scores on tcgen05, softmax + weighted sum of the S neighbour rows on chip' if prob['table_dtype'] == 'bf16' else
               'the unfused attention chain (a(n) as TF32 / 3 x TF32 projections, softmax, weighted gather+sum: the rows are read twice)')
              if agg == 'attention' else 'the fused gather+mean launch') + ', layer 1 on the (x1, x2) pair (B*25 parents, S=10)'
    return {'bound': 'hbm', 'kernel': kernel, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
            'traffic': traffic, 'peak_source': src, 'launches': int(n), 'algorithmic_bytes_per_launch_avg': byt / n,
            'avg_launch_ms': ms / n, 'timed_in': label_timed_in}


def trainable(prob, dtype, tf32):
    import torch
    return (prob['aggregator'] == 'mean' and (prob['prep'] == 'identity' or (prob['prep'] == 'node_embedding' and dtype == torch.float32 and not tf32))) or \
           (prob['aggregator'] in ('max_pool', 'mean_pool') and dtype == torch.bfloat16) or \
           (prob['aggregator'] == 'attention' and prob['prep'] == 'identity' and dtype == torch.bfloat16)


def run_leg(args, workload, B, steps, warmup, main):
    """One workload on this rank's GPU: device-timed steps, (main leg only) the end-to-end host-buffer leg, the train-step leg.
    Returns the leg's result dict on every rank (timings are max over ranks)."""
    import gc
    import numpy as np
    import torch
    import torch.distributed as dist
    import pytorch_graphsage_b200 as g
    from pytorch_graphsage_b200 import synth

    prob = make_problem(workload, args.scale, host_feats=main)
    dtype = torch.bfloat16 if prob['table_dtype'] == 'bf16' else torch.float32
    tf32 = prob['table_dtype'] == 'tf32'
    graph = g.GraphCSR.from_synth(prob['adj'])
    if prob['feats_on_device']:
        table = device_table(prob, dtype)
    else:
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

    ahead = not args.no_ahead
    step_no = [0]

    def step():
        i = step_no[0]
        step_no[0] += 1
        # batch i+1 is drawn while batch i aggregates: its ids exist before this forward is queued (next_ids)
        return model(dev_ids[i % n_batches], table, next_ids=dev_ids[(i + 1) % n_batches] if ahead else None)

    if ahead:
        model.sample_ahead(dev_ids[0], table)
    for i in range(max(3, warmup)):
        step()
    g.default_rng().check()
    graph.check()
    model.profile(True)
    clocks = None
    if main:
        try:
            uuid = 'GPU-' + str(torch.cuda.get_device_properties(LOCAL_RANK).uuid)
        except Exception:
            uuid = str(LOCAL_RANK)
        clocks = ClockSampler(uuid, LOCAL_RANK)
        clocks.prepare()
    barrier()
    if clocks:
        clocks.start()
    launches0 = g.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        step()
    ev1.record()
    barrier()
    launches = g.launch_count() - launches0
    clk = clocks.stop() if clocks else None
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    prof = model.profile_read()
    g.default_rng().check()
    ms_step = ms_total / steps
    value = WORLD * B * ROWS_PER_SEED / (ms_step / 1e3)

    # the dominant kernel alone: a few steps WITHOUT sample-ahead, so that no sampler kernel shares the SMs with it
    # (in the pipelined run above its launches overlap the next batch's draws: the step gets shorter, the kernel longer)
    prof_iso = None
    if ahead:
        model(dev_ids[step_no[0] % n_batches], table)                          # consume the batch the last timed step drew ahead
        step_no[0] += 1
        torch.cuda.synchronize()
        model.profile_read()
        for i in range(10):
            model(dev_ids[(step_no[0] + i) % n_batches], table)
        torch.cuda.synchronize()
        prof_iso = model.profile_read()
        step_no[0] += 10
    model.profile(False)

    res = {'workload': workload, 'value': value, 'unit': UNIT, 'ms_per_step': ms_step, 'steps': steps, 'seeds_per_s': value / ROWS_PER_SEED,
           'dtype': 'bf16' if dtype == torch.bfloat16 else ('tf32' if tf32 else 'f32'),
           'config': workload_config(workload, prob, B, WORLD, args.scale, ahead), 'gpu_launches': int(launches) * WORLD,
           'n_gpus': WORLD}
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj.get('%s:%d' % (workload, B))
        traffic_src = 'static: one `ncu --set full` capture of this kernel on this shape (profiles/traffic.json), not measured in this run'
    timed_in = ('the pipelined steps above (its launches share the SMs with the next batch\'s sampling kernels)' if ahead
                else 'the timed steps above')
    res['roofline'] = dominant_roofline(prob, prof, timed_in, traffic)
    if res['roofline'] is not None:
        res['roofline']['traffic_source'] = traffic_src
        if prof_iso is not None and prof_iso['reduce'][1]:
            iso = dominant_roofline(prob, prof_iso, 'n/a')
            res['roofline']['isolated'] = {'achieved': iso['achieved'], 'frac': iso['frac'], 'launches': iso['launches'],
                                           'avg_launch_ms': iso['avg_launch_ms'],
                                           'note': 'same kernel, same inputs, 10 steps without sample-ahead (nothing else on the SMs)'}
    per = lambda k: prof[k][0] / steps
    res['breakdown_ms_per_step'] = {
        'forward': per('forward'), 'wait_for_sampled_batch': per('wait'), 'layer1_x0x1': per('app0'), 'dominant_aggregate': per('reduce'),
        'project': per('project'), 'layer2': per('layer2'), 'normalize_classifier': per('head'),
        'sample_on_sampler_stream': per('sample'),
        'project_tflops': (prof['project'][3] / 1e12) / (prof['project'][0] / 1e3) if prof['project'][0] > 0 else None,
        'project_gbs_algorithmic': (prof['project'][2] / 1e9) / (prof['project'][0] / 1e3) if prof['project'][0] > 0 else None}
    # whole-step HBM roofline (algorithmic bytes per seed of the mean path)
    if prob['aggregator'] == 'mean':
        e = 2 if dtype == torch.bfloat16 else 4
        d = prob['feats_dim'] or 64
        per_seed = 276 * d * e + 26 * 16 + 275 * 8 + 275 * 16 + 26 * 2 * 128 * e + 27 * 2 * 128 * e
        peak, _ = measured_peaks()
        gbs = per_seed * B / (ms_step / 1e3) / 1e9                              # per GPU (every rank runs B seeds per step)
        res['step_roofline'] = {'bound': 'hbm', 'algorithmic_bytes_per_seed': per_seed, 'achieved': gbs, 'peak': peak, 'unit': 'GB/s',
                                'frac': gbs / peak, 'note': 'the WHOLE step against the HBM roofline (SURVEY.md 8d bytes per seed), per GPU'}
    if clk is not None:
        res['clocks'] = clk

    # ---- end to end through the host-buffer entry (main leg) ------------------------------------------------------
    if main:
        host_ids = [torch.from_numpy(synth.seed_batch(prob, B, seed=17 * RANK + i)).pin_memory() for i in range(n_batches)]
        host_out = torch.empty((B, prob['n_classes']), dtype=torch.float32).pin_memory()

        def host_step(i):                                                          # synchronises every step (D2H of the logits)
            nxt = host_ids[(i + 1) % n_batches] if ahead else None
            model.forward_host(host_ids[i % n_batches], table, host_out, next_ids_host=nxt)

        for i in range(3):
            host_step(i)
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for i in range(3, 3 + steps):
            host_step(i)
        ev1.record()
        barrier()
        if ahead:
            model.forward_host(host_ids[(3 + steps) % n_batches], table, host_out)   # drain the pending batch
        e2e_ms = max_over_ranks(max(ev0.elapsed_time(ev1), 0.0)) / steps
        e2e_wall_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / steps
        e2e_ms = max(e2e_ms, e2e_wall_ms)
        res['e2e'] = {'value': WORLD * B * ROWS_PER_SEED / (e2e_ms / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': 8 * B,
                      'd2h_bytes_per_step': 4 * B * prob['n_classes'], 'ms_per_step': e2e_ms}

    # ---- one optimiser step per batch (forward + loss + backward + gradient all-reduce + clip + Adam) ---------------
    if trainable(prob, dtype, tf32) and not args.no_train:
        from torch.nn import functional as F
        if prob['task'] == 'regression_mae':                                       # problem.py:39-41: l1 loss on (B, 1) predictions
            tgt_all = torch.from_numpy(prob['targets']).cuda()
            loss_fn = lambda preds, t: F.l1_loss(preds, t.view_as(preds))
        else:
            tgt_all = torch.from_numpy(prob['targets'].reshape(-1)).cuda()
            loss_fn = F.cross_entropy
        tgts = [tgt_all[i] for i in dev_ids]
        # clip_grad_norm 5 + Adam as one native call over flat buffers (the model's own parallel.FusedAdam); --torch-adam: the stock pair
        opt = torch.optim.Adam(model.parameters(), lr=0.01) if args.torch_adam else None
        side = torch.cuda.Stream()
        k_train = max(3, min(steps, 30))

        def train_step(i):
            model.train_step(dev_ids[i % n_batches], table, tgts[i % n_batches], loss_fn, optimizer=opt, grad_scale=1.0 / WORLD,
                             overlap_stream=side, next_ids=dev_ids[(i + 1) % n_batches] if ahead else None)

        if ahead and model.optimizer.flat is None and opt is None:
            model.optimizer._materialize()                                         # (re-creates the engines: before the first sample-ahead)
        if ahead:
            model.sample_ahead(dev_ids[0], table)
        for i in range(3):
            train_step(i)
        barrier()
        prof_train = os.environ.get('GSAGE_BENCH_PROFILE_TRAIN') == '1'           # ncu --profile-from-start off: the train steps only
        if prof_train:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        ev0.record()
        for i in range(3, 3 + k_train):
            train_step(i)
        ev1.record()
        barrier()
        if prof_train:
            torch.cuda.profiler.stop()
        if ahead:
            model(dev_ids[(3 + k_train) % n_batches], table)                       # drain the pending batch
        t_ms = max_over_ranks(ev0.elapsed_time(ev1)) / k_train
        res['train'] = {'ms_per_step': t_ms, 'seeds_per_s': WORLD * B / (t_ms / 1e3), 'value': WORLD * B * ROWS_PER_SEED / (t_ms / 1e3), 'steps': k_train,
                        'allreduce_bytes_per_step': int(model._bucket().flat.numel()) * 4,
                        'collective': ('one flat fp32 gradient bucket over %d ranks: %s' % (WORLD, model._bucket().collective)) if WORLD > 1 else 'none (1 GPU)',
                        'note': 'loss is stock torch; clip + Adam are ' + ('stock torch' if args.torch_adam else 'one native call (gsage_adam_step)')}
    model.check()
    del model, table, graph, dev_ids
    gc.collect()
    torch.cuda.empty_cache()
    res['_prob'] = prob
    return res


def sharded_parity_check():
    """N > 1 only: the seed-sharded train step against the unsharded one, on every rank of the job.  Model A: each rank
    runs ITS slice of one global batch (shard=(global, first)) and the gradients are all-reduced with local/global weights.
    Model B: every rank runs the WHOLE batch (weight 1/N each, so the all-reduce returns the full-batch gradient).  Checked on
    every rank: the sampled ids of the slice equal the matching slice of the full run bit for bit (same MT19937 stream positions),
    logits agree, all-reduced gradients agree.  (Single-GPU parity against the reference is what tests/ establishes.)"""
    import numpy as np
    import torch
    import torch.distributed as dist
    from torch.nn import functional as F
    import pytorch_graphsage_b200 as g
    from pytorch_graphsage_b200 import synth
    from pytorch_graphsage_b200.parallel import shard_seeds
    prob = synth.make_problem('tiny', seed=1)
    graph = g.GraphCSR.from_synth(prob['adj'])
    table = g.FeatureTable(prob['feats'], torch.float32)
    ids = torch.from_numpy(synth.seed_batch(prob, 101, seed=9))                 # 101 seeds: uneven shards at every N
    targets = torch.from_numpy(prob['targets'].reshape(-1))[ids]
    S1, S2 = FANOUT

    def build():
        torch.manual_seed(77)
        return g.GSSupervised(input_dim=prob['feats_dim'], n_nodes=prob['n_nodes'], n_classes=prob['n_classes'], layer_specs=layer_specs(),
                              aggregator_class=g.aggregator_lookup['mean'], prep_class=g.prep_lookup['identity'],
                              sampler_class=g.sampler_lookup['sparse_uniform_neighbor_sampler'], adj=graph, train_adj=graph).cuda()

    mine, tmine = shard_seeds(ids, RANK, WORLD), shard_seeds(targets, RANK, WORLD)
    first = int(sum(shard_seeds(ids, r, WORLD).shape[0] for r in range(RANK)))
    a = build()
    g.set_seeds(4242)
    pa = a.train_step(mine, table, tmine.cuda(), F.cross_entropy, optimizer=False, clip=None, grad_scale=mine.shape[0] / ids.shape[0],
                      shard=(ids.shape[0], first))
    ids2_a = a.peek('ids2').clone()
    ga = {n: p.grad.clone() for n, p in a.named_parameters()}
    b = build()
    g.set_seeds(4242)
    pb = b.train_step(ids, table, targets.cuda(), F.cross_entropy, optimizer=False, clip=None, grad_scale=1.0 / WORLD)
    ids2_b = b.peek('ids2')
    lo, n_mine = first * S1 * S2, mine.shape[0] * S1 * S2
    ids_ok = bool(torch.equal(ids2_a, ids2_b[lo:lo + n_mine]))
    logit_diff = float((pa - pb[first:first + mine.shape[0]]).abs().max())
    grad_diff = max(float((ga[n] - p.grad).abs().max() / (p.grad.abs().max() + 1e-12)) for n, p in b.named_parameters())
    t = torch.tensor([0.0 if ids_ok else 1.0, logit_diff, grad_diff], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    bad_ids, logit_diff, grad_diff = [float(x) for x in t.tolist()]
    return {'ranks': WORLD, 'global_batch': int(ids.shape[0]), 'ids_bit_exact_on_every_rank': bad_ids == 0.0, 'max_abs_logit_diff': logit_diff,
            'max_rel_grad_diff': grad_diff, 'ok': bad_ids == 0.0 and logit_diff < 1e-4 and grad_diff < 2e-3,
            'collective': a._bucket().collective,
            'what': 'seed-sharded train step (uneven shards, local/global weights, gradient all-reduce) vs the unsharded step on the same '
                    'MT19937 stream, checked on every rank'}


def parse_legs(args):
    if args.workload != 'reddit' or args.legs == 'none':
        return []
    if args.legs == 'default':
        return list(DEFAULT_LEGS)
    legs = []
    for item in args.legs.split(','):
        f = item.split(':')
        legs.append((f[0], int(f[1]) if len(f) > 1 else 16384, int(f[2]) if len(f) > 2 else 40))
    return legs


def run_ours(args):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(LOCAL_RANK)                                         # one process per GPU
    if WORLD > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() not in ('INFO', 'TRACE'):
            os.environ['NCCL_DEBUG'] = 'WARN'                                  # keep stdout to the one JSON line
        dist.init_process_group('nccl', device_id=torch.device('cuda', LOCAL_RANK))

    main = run_leg(args, args.workload, args.batch, args.steps, args.warmup, main=True)
    prob = main.pop('_prob')
    legs = {}
    for workload, B, steps in parse_legs(args):
        t0 = time.time()
        try:
            leg = run_leg(args, workload, B, steps, 3, main=False)
            leg.pop('_prob')
            leg['wall_s_incl_setup'] = time.time() - t0
        except Exception as exc:                                               # a leg that cannot run must not take the headline with it
            if WORLD > 1:
                raise
            leg = {'workload': workload, 'error': '%s: %s' % (type(exc).__name__, exc)}
        legs[workload] = leg

    parity = sharded_parity_check() if WORLD > 1 else None

    if RANK != 0:
        if WORLD > 1:
            dist.destroy_process_group()
        return

    line = {
        'metric': METRIC, 'value': main['value'], 'unit': UNIT, 'n_gpus': WORLD, 'steps': args.steps, 'warmup': max(3, args.warmup),
        'ms_per_step': main['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': main['dtype'], 'data': 'synthetic', 'config': main['config'], 'clocks': main.get('clocks'),
        'e2e': main['e2e'], 'gpu_launches': main['gpu_launches'], 'seeds_per_s': main['seeds_per_s'],
        'roofline': main['roofline'], 'breakdown_ms_per_step': main['breakdown_ms_per_step'],
    }
    if 'step_roofline' in main:
        line['step_roofline'] = main['step_roofline']
    if 'train' in main:
        line['train'] = main['train']
    if legs:
        line['configs'] = legs
    if parity is not None:
        line['multi_gpu_parity'] = parity
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
