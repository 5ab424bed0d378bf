"""
Synthetic problems of the shapes BASELINE.json names (there is no network for Reddit / Pokec / Cora).

Everything follows the reference's on-disk conventions for the *sparse* problem files
(/root/reference/utils/convert.py:100-126, :128-131, :209-211):

  * node i is row i+1 of the adjacency; row 0 is the empty dummy node;
  * row i+1 stores the neighbours of i as VALUES `nbr+1` in COLUMNS 0..deg-1;
  * the file holds the 3 x nnz array [v; r; c] and the loader rebuilds `csr_matrix((v, (r, c)))`
    (/root/reference/problem.py:70-72), so shape = (max row + 1, max degree);
  * features / targets get the dummy row FIRST (row 0 = zeros).

Generators are seeded `numpy.random.RandomState`s of their own -- they never touch the global
`np.random` stream that the sampler shares with the reference.
"""

import numpy as np

# name -> (n_nodes, n_edges (directed, stored), feature dim, pareto alpha, degree clip, n_classes)
SHAPES = {
    'cora':   dict(n_nodes=2708,     n_edges=13264,      d=1433, alpha=2.0, clip=168,   n_classes=7),
    'reddit': dict(n_nodes=232965,   n_edges=11000000,   d=602,  alpha=1.5, clip=20000, n_classes=41),
    'pokec':  dict(n_nodes=1632803,  n_edges=30622564,   d=64,   alpha=1.8, clip=15000, n_classes=1),
    'plaw2m': dict(n_nodes=2000000,  n_edges=40000000,   d=256,  alpha=1.2, clip=50000, n_classes=41),
    'big10m': dict(n_nodes=10000000, n_edges=200000000,  d=256,  alpha=1.5, clip=50000, n_classes=41),
    'tiny':   dict(n_nodes=1000,     n_edges=12000,      d=48,   alpha=1.5, clip=200,   n_classes=5),
}


def powerlaw_degrees(n_nodes, n_edges, alpha, clip, rs):
    """Pareto(alpha) degrees rescaled so they sum to ~n_edges, clipped to [0, clip]; the LAST node
    always keeps >= 1 edge so the inferred CSR shape covers every id."""
    raw = rs.pareto(alpha, size=n_nodes) + 1.0
    deg = np.minimum(np.floor(raw * (n_edges / raw.sum())), clip).astype(np.int64)
    # top up round-off / clipping losses uniformly
    short = n_edges - int(deg.sum())
    if short > 0:
        bump = rs.randint(0, n_nodes, size=short)
        np.add.at(deg, bump, 1)
        deg = np.minimum(deg, clip)
    deg[-1] = max(deg[-1], 1)
    return deg


def make_sparse_adjacency(n_nodes, n_edges, alpha=1.5, clip=20000, seed=0, isolated_frac=0.0):
    """Returns a dict with the device-ready CSR (`indptr`, `data`), the reference's `(v, r, c)`
    triplets view helpers, and `shape`.  `data[indptr[r] + k]` is the k-th neighbour (+1) of row r."""
    rs = np.random.RandomState(seed)
    deg = powerlaw_degrees(n_nodes, n_edges, alpha, clip, rs)
    if isolated_frac > 0:
        kill = rs.rand(n_nodes) < isolated_frac
        kill[-1] = False
        deg[kill] = 0
    indptr = np.zeros(n_nodes + 2, dtype=np.int64)        # +1 dummy row, +1 fence
    np.cumsum(deg, out=indptr[2:])
    nnz = int(indptr[-1])
    data = rs.randint(1, n_nodes + 1, size=nnz).astype(np.int64)   # neighbour ids in the +1 space
    return dict(indptr=indptr, data=data, shape=(n_nodes + 1, int(deg.max())), n_nodes=n_nodes, nnz=nnz)


def triplets(adj):
    """The 3 x nnz `[v; r; c]` array the reference's problem files hold (utils/convert.py:128-131)."""
    indptr, data = adj['indptr'], adj['data']
    deg = np.diff(indptr)
    r = np.repeat(np.arange(indptr.shape[0] - 1, dtype=np.int64), deg)
    c = np.arange(data.shape[0], dtype=np.int64) - np.repeat(indptr[:-1], deg)
    return np.vstack([data, r, c])


def make_features(n_nodes, d, seed=0, dtype=np.float32, kind='normal'):
    """(n_nodes + 1, d) table, dummy row 0 = zeros.  'normal' ~ N(0,1); 'bow' = row-normalised
    binary bag-of-words (Cora-like, /root/reference/utils/convert-cora.py:63)."""
    rs = np.random.RandomState(seed + 7919)
    if kind == 'bow':
        f = (rs.rand(n_nodes + 1, d) < 0.0127).astype(np.float32)
        f[np.arange(n_nodes + 1), rs.randint(0, d, n_nodes + 1)] = 1.0
        f /= f.sum(axis=1, keepdims=True)
    else:
        f = np.empty((n_nodes + 1, d), dtype=np.float32)
        step = max(1, (1 << 22) // max(d, 1))
        for lo in range(0, n_nodes + 1, step):               # chunked: 10 M x 256 does not fit twice
            hi = min(n_nodes + 1, lo + step)
            f[lo:hi] = rs.standard_normal((hi - lo, d)).astype(np.float32)
    f[0] = 0.0
    return f.astype(dtype, copy=False)


def make_problem(name, seed=0, with_feats=True, scale=1.0):
    """A stand-in for the fields of `NodeProblem` the hot path reads (/root/reference/problem.py:80-106)."""
    s = dict(SHAPES[name])
    n_nodes = max(16, int(s['n_nodes'] * scale))
    n_edges = max(64, int(s['n_edges'] * scale))
    adj = make_sparse_adjacency(n_nodes, n_edges, s['alpha'], s['clip'], seed)
    feats = None
    if with_feats:
        feats = make_features(n_nodes, s['d'], seed, kind='bow' if name == 'cora' else 'normal')
    rs = np.random.RandomState(seed + 104729)
    if s['n_classes'] > 1:
        targets = rs.randint(0, s['n_classes'], size=(n_nodes + 1, 1))
        task = 'classification'
    else:
        targets = rs.randint(15, 61, size=(n_nodes + 1, 1)).astype(np.float32)
        task = 'regression_mae'
    return dict(name=name, adj=adj, train_adj=adj, feats=feats, feats_dim=(s['d'] if with_feats else None),
                n_nodes=n_nodes + 1, n_classes=s['n_classes'], targets=targets, task=task)


def seed_batch(problem, batch, seed=0):
    """A batch of seed ids in the reference's id space (1..N; never the dummy)."""
    rs = np.random.RandomState(seed + 15485863)
    return rs.randint(1, problem['n_nodes'], size=batch).astype(np.int64)
