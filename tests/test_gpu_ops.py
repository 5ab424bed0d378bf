"""Kernel-level parity of the C-ABI ops against plain torch fp32 / fp64 on the same inputs.  GPU only.

Tolerances (stated, per op):
  gather_reduce fp32 table        rtol 1e-5  atol 1e-6   (fp32 sums of <= 40 terms, different association)
  gather_reduce bf16 table        exact inputs (bf16 -> fp32 is exact), fp32 accumulate: same as above vs the
                                  upcast table; bf16 OUTPUT adds one rounding: rtol 8e-3 (2^-7)
  linear fp32 (FFMA kernel)       rtol 1e-4  atol 1e-5 vs fp64 (k up to 1433 products)
"""
import numpy as np
import pytest
import torch
from torch.nn import functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def g():
    import pytorch_graphsage_b200 as g
    return g


def _table(rows, d, dtype, seed=0):
    gen = torch.Generator().manual_seed(seed)
    t = torch.randn((rows, d), generator=gen)
    t[0] = 0
    if dtype == torch.bfloat16:
        t = t.to(torch.bfloat16).float()          # values exactly representable in bf16
    return t


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('d', [3, 20, 64, 100, 256, 602, 1433])
@pytest.mark.parametrize('S', [1, 10, 25, 40])
def test_gather_mean(g, dtype, d, S):
    rows, n = 977, 301
    host = _table(rows, d, dtype)
    store, _ = g.ops.pad_table(host, dtype)
    table = store[:, :d]
    ids = torch.randint(0, rows, (n * S,), generator=torch.Generator().manual_seed(d * S))
    want = host[ids].double().view(n, S, d).mean(dim=1)
    got = g.ops.gather_reduce(table, ids.cuda(), n, S, 'mean', out_dtype=torch.float32)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6)
    if dtype == torch.bfloat16:
        got16 = g.ops.gather_reduce(table, ids.cuda(), n, S, 'mean', out_dtype=torch.bfloat16)
        np.testing.assert_allclose(got16.float().cpu().numpy(), want.numpy(), rtol=8e-3, atol=1e-3)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('d', [32, 512])
def test_contiguous_reduce_max_and_mean(g, dtype, d):
    """ids=None: the `neibs.view(N, S, d)` case (pool aggregators reduce the MLP output)."""
    n, S = 130, 10
    host = _table(n * S, d, dtype, seed=5)
    h = g.ops.pad_table(host, dtype)[0][:, :d]
    got_max = g.ops.gather_reduce(h, None, n, S, 'max', out_dtype=torch.float32)
    got_mean = g.ops.gather_reduce(h, None, n, S, 'mean', out_dtype=torch.float32)
    assert torch.equal(got_max.cpu(), host.view(n, S, d).max(dim=1)[0])          # max is exact
    np.testing.assert_allclose(got_mean.cpu().numpy(), host.double().view(n, S, d).mean(dim=1).numpy(), rtol=1e-5, atol=1e-6)


def test_weighted_sum(g):
    rows, n, S, d = 500, 77, 25, 48
    host = _table(rows, d, torch.float32, seed=9)
    ids = torch.randint(0, rows, (n * S,))
    w = torch.softmax(torch.randn(n, S), dim=1)
    want = (host[ids].double().view(n, S, d) * w.double().unsqueeze(-1)).sum(dim=1)
    got = g.ops.gather_reduce(g.ops.pad_table(host)[0][:, :d], ids.cuda(), n, S, 'sum', weights=w.view(-1).cuda().contiguous())
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6)


def test_gather_rows_and_unaligned_input(g):
    host = _table(300, 602, torch.float32, seed=2)              # 2408-byte rows: not 16-byte aligned -> padded copy
    ids = torch.randint(0, 300, (1000,))
    got = g.ops.gather_rows(host.cuda(), ids.cuda())
    assert torch.equal(got.cpu(), host[ids])


def test_out_of_range_ids_read_as_zero_rows(g):
    host = _table(10, 8, torch.float32)
    got = g.ops.gather_reduce(g.ops.pad_table(host)[0][:, :8], torch.tensor([1, 99, 2, -5]).cuda(), 2, 2, 'sum')
    assert torch.equal(got.cpu(), torch.stack([host[1], host[2]]))


@pytest.mark.parametrize('a_dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('n,d,O', [(1, 7, 5), (64, 64, 128), (301, 602, 128), (130, 1433, 16), (1000, 256, 41), (77, 512, 130)])
def test_linear(g, a_dtype, n, d, O):
    gen = torch.Generator().manual_seed(n + d + O)
    a = _table(n + 40, d, a_dtype, seed=n)
    w = torch.randn((O, d), generator=gen) / d ** 0.5
    b = torch.randn((O,), generator=gen)
    ids = torch.randint(0, n + 40, (n,), generator=gen)
    a_dev = g.ops.pad_table(a, a_dtype)[0][:, :d]
    for act, fn in ((None, lambda t: t), ('relu', torch.relu), ('tanh', torch.tanh)):
        want = fn(a[ids].double() @ w.double().t() + b.double())
        got = g.ops.linear([dict(a=a_dev, ids=ids.cuda(), w=w.cuda(), bias=b.cuda())], n, act=act)
        np.testing.assert_allclose(got.detach().cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)
    want = a[:n].double() @ w.double().t()                       # no gather, no bias
    got = g.ops.linear([dict(a=a_dev, w=w.cuda())], n)
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)


def test_linear_two_segments_concat(g):
    """[fc_x(x) | fc_neib(m)] in one launch (nn_modules.py:200)."""
    n, d, O = 257, 100, 24
    x, m = torch.randn(n, d), torch.randn(n, d)
    wx, wn = torch.randn(O, d) / 10, torch.randn(O, d) / 10
    want = torch.relu(torch.cat([x.double() @ wx.double().t(), m.double() @ wn.double().t()], dim=1))
    got = g.ops.linear([dict(a=g.ops.aligned_rows(x.cuda()), w=wx.cuda(), col0=0),
                        dict(a=g.ops.aligned_rows(m.cuda()), w=wn.cuda(), col0=O)], n, act='relu')
    assert got.shape == (n, 2 * O)
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)


def test_attention_weights_and_l2norm(g):
    n, S, H = 99, 10, 32
    na, xa = torch.randn(n * S, H), torch.randn(n, H)
    want = torch.softmax(torch.einsum('nsh,nh->ns', na.double().view(n, S, H), xa.double()), dim=1)
    got = g.ops.attention_weights(na.cuda(), xa.cuda(), n, S)
    np.testing.assert_allclose(got.view(n, S).cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-7)
    z = torch.randn(50, 256)
    z[3] = 0                                                        # F.normalize eps path
    got = g.ops.l2_normalize(z.cuda())
    np.testing.assert_allclose(got.cpu().numpy(), torch.nn.functional.normalize(z.double(), dim=1).numpy(), rtol=1e-5, atol=1e-7)
    with pytest.raises(ValueError):
        g.ops.attention_weights(na.cuda(), xa.cuda(), n * S, 1)     # S == 1 is out of contract


# ---- tcgen05 projection kernel (bf16 x bf16 -> fp32 accumulate in TMEM) -------------------------------------
# Tolerance: operands are exactly representable in bf16, products are exact in fp32, only the accumulation
# order differs from fp64: rtol 2e-4 / atol 2e-4 for fp32 output; bf16 output adds one rounding (2^-8 rel).

def _bf16(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize('n,d,O', [(128, 64, 128), (1, 64, 16), (129, 100, 128), (1000, 602, 128), (4097, 256, 256),
                                   (300, 1433, 32), (640, 512, 48)])
@pytest.mark.parametrize('out_dtype', [torch.float32, torch.bfloat16])
def test_umma_linear_single_segment(g, n, d, O, out_dtype):
    gen = torch.Generator().manual_seed(n * 7 + d + O)
    rows = n + 50
    a = _bf16(torch.randn((rows, d), generator=gen))
    w = _bf16(torch.randn((O, d), generator=gen) / d ** 0.5)
    b = torch.randn((O,), generator=gen)
    ids = torch.randint(0, rows, (n,), generator=gen)
    a_dev = g.ops.pad_table(a.float(), torch.bfloat16)[0][:, :d]
    w_dev = g.ops.pad_table(w.float(), torch.bfloat16)[0][:, :d]
    tol = dict(rtol=2e-4, atol=2e-4) if out_dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    for act, fn in ((None, lambda t: t), ('relu', torch.relu)):
        want = fn(a[ids].double() @ w.double().t() + b.double())
        got = g.ops.linear([dict(a=a_dev, ids=ids.cuda(), w=w_dev, bias=b.cuda())], n, act=act, out_dtype=out_dtype, exact=False)
        np.testing.assert_allclose(got.float().cpu().numpy(), want.numpy(), **tol)
    want = a[:n].double() @ w.double().t()
    got = g.ops.linear([dict(a=a_dev, w=w_dev)], n, out_dtype=out_dtype, exact=False)
    np.testing.assert_allclose(got.float().cpu().numpy(), want.numpy(), **tol)


@pytest.mark.parametrize('n,d,O', [(777, 602, 128), (128 * 150 + 3, 64, 64), (50, 256, 128)])
def test_umma_concat_with_self(g, n, d, O):
    """[fc_x(table[ids]) | fc_neib(m)] + relu in one tensor-core launch (two TMEM accumulators per tile)."""
    gen = torch.Generator().manual_seed(n)
    table = _bf16(torch.randn((n + 99, d), generator=gen))
    m = _bf16(torch.randn((n, d), generator=gen))
    wx, wn = _bf16(torch.randn((O, d), generator=gen) / d ** 0.5), _bf16(torch.randn((O, d), generator=gen) / d ** 0.5)
    ids = torch.randint(0, n + 99, (n,), generator=gen)
    pad = lambda t: g.ops.pad_table(t.float(), torch.bfloat16)[0][:, :t.shape[1]]
    want = torch.relu(torch.cat([table[ids].double() @ wx.double().t(), m.double() @ wn.double().t()], dim=1))
    got = g.ops.linear([dict(a=pad(table), ids=ids.cuda(), w=pad(wx), col0=0), dict(a=pad(m), w=pad(wn), col0=O)], n,
                       act='relu', out_dtype=torch.float32, exact=False)
    assert got.shape == (n, 2 * O)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-4, atol=2e-4)


def test_umma_wide_output_is_split(g):
    """O = 512 (the pool MLP) does not fit one 256-column accumulator: the dispatcher issues column blocks."""
    n, d, O = 1500, 64, 512
    gen = torch.Generator().manual_seed(3)
    a, w, b = _bf16(torch.randn((n, d), generator=gen)), _bf16(torch.randn((O, d), generator=gen) / 8), torch.randn((O,), generator=gen)
    pad = lambda t: g.ops.pad_table(t.float(), torch.bfloat16)[0][:, :t.shape[1]]
    want = torch.relu(a.double() @ w.double().t() + b.double())
    got = g.ops.linear([dict(a=pad(a), w=pad(w), bias=b.cuda())], n, act='relu', exact=False)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize('n,d,O,S', [(300, 602, 128, 10), (129, 64, 64, 25), (1000, 256, 128, 3), (50, 100, 16, 40)])
def test_fused_gather_mean_projection(g, n, d, O, S):
    """[fc_x(table[ids]) | fc_neib(mean_j table[nb_ids])]: the gather+mean happens inside the projection kernel's
    operand load (tcgen05 path, bf16) -- compared with fp64 on the same bf16 operands, and with the fp32 FFMA path."""
    gen = torch.Generator().manual_seed(n + S)
    rows = 2000
    table = _bf16(torch.randn((rows, d), generator=gen))
    wx, wn = _bf16(torch.randn((O, d), generator=gen) / d ** 0.5), _bf16(torch.randn((O, d), generator=gen) / d ** 0.5)
    ids = torch.randint(0, rows, (n,), generator=gen)
    nb = torch.randint(0, rows, (n * S,), generator=gen)
    pad = lambda t: g.ops.pad_table(t.float(), torch.bfloat16)[0][:, :t.shape[1]]
    mean = table[nb].float().view(n, S, d).sum(dim=1) * (1.0 / S)               # fp32 sum in j order, like the kernel
    mean_bf16 = mean.to(torch.bfloat16)                                         # the kernel rounds the mean to bf16 in smem
    want = torch.relu(torch.cat([table[ids].double() @ wx.double().t(), mean_bf16.double() @ wn.double().t()], dim=1))
    segs = [dict(a=pad(table), ids=ids.cuda(), w=pad(wx), col0=0), dict(a=pad(table), ids=nb.cuda(), w=pad(wn), col0=O, S=S)]
    got = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact=False)
    # bf16 rounding of the mean can flip by one ulp vs the torch emulation (summation order): 2^-8 relative on that operand
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-3, atol=2e-3)
    exact = torch.relu(torch.cat([table[ids].double() @ wx.double().t(),
                                  table[nb].double().view(n, S, d).mean(dim=1) @ wn.double().t()], dim=1))
    got32 = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact=True)   # FFMA path: mean kept in fp32
    np.testing.assert_allclose(got32.cpu().numpy(), exact.numpy(), rtol=1e-4, atol=1e-5)


def test_fused_contiguous_mean_projection(g):
    """ids=None, S>1: the layer-2 case -- neighbours are rows r*S+j of the previous layer's output."""
    n, d, O, S = 200, 256, 128, 25
    gen = torch.Generator().manual_seed(8)
    h = _bf16(torch.randn((n + n * S, d), generator=gen))
    wx, wn = _bf16(torch.randn((O, d), generator=gen) / 16), _bf16(torch.randn((O, d), generator=gen) / 16)
    pad = lambda t: g.ops.pad_table(t.float(), torch.bfloat16)[0][:, :t.shape[1]]
    hd = pad(h)
    mean = (h[n:].float().view(n, S, d).sum(dim=1) * (1.0 / S)).to(torch.bfloat16)
    want = torch.cat([h[:n].double() @ wx.double().t(), mean.double() @ wn.double().t()], dim=1)
    got = g.ops.linear([dict(a=hd[:n], w=pad(wx), col0=0), dict(a=hd[n:], w=pad(wn), col0=O, S=S)], n, exact=False)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize('n,d,O', [(300, 64, 128), (1000, 256, 64), (129, 100, 16)])
def test_umma_tf32_projection(g, n, d, O):
    """fp32 operands on the tensor cores as TF32 (exact=False): products carry a 10-bit mantissa, fp32 accumulate.
    Stated tolerance: 2e-3 * sqrt(d) absolute on unit-variance operands (measured errors are ~10x smaller)."""
    gen = torch.Generator().manual_seed(d)
    table = torch.randn((n + 77, d), generator=gen)
    m = torch.randn((n, d), generator=gen)
    wx, wn = torch.randn((O, d), generator=gen) / d ** 0.5, torch.randn((O, d), generator=gen) / d ** 0.5
    ids = torch.randint(0, n + 77, (n,), generator=gen)
    pad = lambda t: g.ops.pad_table(t, torch.float32)[0][:, :t.shape[1]]
    segs = [dict(a=pad(table), ids=ids.cuda(), w=pad(wx), col0=0), dict(a=pad(m), w=pad(wn), col0=O)]
    want = torch.relu(torch.cat([table[ids].double() @ wx.double().t(), m.double() @ wn.double().t()], dim=1))
    got = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact=False)
    err = (got.cpu().double() - want).abs().max().item()
    assert err < 2e-3 * d ** 0.5, err
    assert err > 0 or d < 8                       # it really ran in reduced precision (the FFMA path would be ~1e-6)
    exact = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact=True)
    np.testing.assert_allclose(exact.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('n,H,h_dtype', [(1, 8, torch.float32), (37, 64, torch.float32), (1000, 512, torch.float32), (129, 64, torch.bfloat16)])
def test_lstm_cell_equals_torch(g, n, H, h_dtype):
    """gsage_lstm_cell against torch.nn.LSTMCell's arithmetic (gate order i, f, g, o), three chained steps from the zero state."""
    gen = torch.Generator().manual_seed(n + H)
    b_ih, b_hh = torch.randn((4 * H,), generator=gen), torch.randn((4 * H,), generator=gen)
    c_ref, h_ref = torch.zeros((n, H), dtype=torch.float64), torch.zeros((n, H), dtype=torch.float64)
    c = torch.full((n, H), 7.0, device='cuda')                          # garbage: the first step must not read it
    h = torch.full((n, H), 7.0, device='cuda').to(h_dtype)
    for t in range(3):
        gx, gh = torch.randn((n, 4 * H), generator=gen), torch.randn((n, 4 * H), generator=gen)
        gates = gx.double() + b_ih.double() + b_hh.double() + (gh.double() if t > 0 else 0.0)
        i, f, gg, o = gates[:, :H], gates[:, H:2 * H], gates[:, 2 * H:3 * H], gates[:, 3 * H:]
        c_ref = torch.sigmoid(f) * c_ref + torch.sigmoid(i) * torch.tanh(gg)
        h_ref = torch.sigmoid(o) * torch.tanh(c_ref)
        g.ops.lstm_cell(gx.cuda(), gh.cuda() if t > 0 else None, b_ih.cuda(), b_hh.cuda(), c, h, first=(t == 0))
        np.testing.assert_allclose(c.cpu().numpy(), c_ref.numpy(), rtol=1e-5, atol=1e-6)
        tol = dict(rtol=1e-5, atol=1e-6) if h_dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
        np.testing.assert_allclose(h.float().cpu().numpy(), h_ref.numpy(), **tol)


@pytest.mark.parametrize('n,S,d,H,gather', [(40, 10, 20, 64, True), (33, 25, 64, 32, False), (5, 3, 7, 8, True)])
def test_lstm_aggregator_equals_torch_lstm(g, n, S, d, H, gather):
    """operators.LSTMAggregator (narrow and id-taking entries) against the stock nn.LSTM it holds its parameters in
    (nn_modules.py:276-279: batch_first, last hidden state)."""
    torch.manual_seed(n + S)
    agg = g.aggregator_lookup['lstm'](input_dim=d, output_dim=16, activation=F.relu, hidden_dim=H)
    table = torch.randn((n * S + 50, d))
    x = torch.randn((n, d))
    with torch.no_grad():
        if gather:
            ids_self = torch.randint(0, table.shape[0], (n,))
            ids_nb = torch.randint(0, table.shape[0], (n * S,))
            xs, nb = table[ids_self], table[ids_nb]
        else:
            xs, nb = x, table[:n * S]
        seq, _ = agg.lstm(nb.view(n, S, d))
        want = F.relu(torch.cat([agg.fc_x(xs), agg.fc_neib(seq[:, -1])], dim=1))
    agg = agg.cuda()
    if gather:
        got = agg.forward_ids(table.cuda(), ids_self.cuda(), ids_nb.cuda(), S)
    else:
        got = agg(xs.cuda(), nb.cuda())
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('n,d,O', [(300, 64, 128), (128 * 200 + 5, 64, 128), (1000, 100, 64), (129, 256, 32), (777, 64, 16)])
def test_three_tf32_projection_has_fp32_accuracy(g, n, d, O, monkeypatch):
    """exact='x3': fp32 operands on the tensor cores as 3 x TF32 (a = a_hi + a_lo, w = w_hi + w_lo, three products, fp32
    accumulate).  Same bar as the FFMA kernel (rtol 1e-4 / atol 1e-5 against float64), and it must be a different kernel
    than both the FFMA one (GSAGE_FP32_FFMA=1) and the single-pass TF32 one (whose error is ~100x larger)."""
    gen = torch.Generator().manual_seed(n + d)
    table = torch.randn((n + 77, d), generator=gen)
    m = torch.randn((n, d), generator=gen)
    wx, wn = torch.randn((O, d), generator=gen) / d ** 0.5, torch.randn((O, d), generator=gen) / d ** 0.5
    bias = torch.randn((O,), generator=gen)
    ids = torch.randint(0, n + 77, (n,), generator=gen)
    pad = lambda t: g.ops.pad_table(t, torch.float32)[0][:, :t.shape[1]]
    segs = [dict(a=pad(table), ids=ids.cuda(), w=pad(wx), col0=0, bias=bias.cuda()), dict(a=pad(m), w=pad(wn), col0=O)]
    want = torch.relu(torch.cat([table[ids].double() @ wx.double().t() + bias.double(), m.double() @ wn.double().t()], dim=1))
    got = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact='x3').cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-4, atol=1e-5)
    err_x3 = (got.double() - want).abs().max().item()
    tf32 = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact=False).cpu()
    err_tf32 = (tf32.double() - want).abs().max().item()
    assert err_x3 < err_tf32 / 20, (err_x3, err_tf32)
    assert err_x3 < 4e-6 * d ** 0.5, err_x3
    monkeypatch.setenv('GSAGE_FP32_FFMA', '1')
    ffma = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact='x3').cpu()
    np.testing.assert_allclose(ffma.numpy(), want.numpy(), rtol=1e-4, atol=1e-5)
    assert not torch.equal(ffma, got)              # different summation order: the tensor-core path really ran


@pytest.mark.parametrize('n,S,d,H', [(100, 10, 64, 512), (37, 25, 64, 512), (513, 3, 100, 32), (8, 128, 64, 48), (3000, 10, 256, 512),
                                     (257, 25, 256, 512), (1000, 10, 602, 512), (50, 64, 64, 200), (64, 2, 72, 136)])
@pytest.mark.parametrize('reduce', ['max', 'mean'])
def test_umma_pooled_epilogue(g, n, S, d, H, reduce):
    """relu(MLP) on tcgen05 with the max / mean over the S neighbour rows of each parent done in the epilogue."""
    gen = torch.Generator().manual_seed(n * S + d)
    rows = 3000
    table = _bf16(torch.randn((rows, d), generator=gen))
    w, b = _bf16(torch.randn((H, d), generator=gen) / d ** 0.5), torch.randn((H,), generator=gen)
    ids = torch.randint(0, rows, (n * S,), generator=gen)
    pad = lambda t: g.ops.pad_table(t.float(), torch.bfloat16)[0][:, :t.shape[1]]
    h = torch.relu(table[ids].double() @ w.double().t() + b.double()).view(n, S, H)
    want = h.max(dim=1)[0] if reduce == 'max' else h.mean(dim=1)
    got = g.ops.linear_pooled(pad(table), pad(w), n, S, reduce, ids=ids.cuda(), bias=b.cuda())
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-4, atol=2e-4)
    # contiguous neighbours (ids=None): rows p*S+j of the operand itself
    nb = _bf16(torch.randn((n * S, d), generator=gen))
    h = torch.relu(nb.double() @ w.double().t() + b.double()).view(n, S, H)
    want = h.max(dim=1)[0] if reduce == 'max' else h.mean(dim=1)
    got = g.ops.linear_pooled(pad(nb), pad(w), n, S, reduce, bias=b.cuda(), out_dtype=torch.bfloat16)
    np.testing.assert_allclose(got.float().cpu().numpy(), want.numpy(), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('n,S,d,H,dtype', [(40003, 10, 64, 512, 'bf16'), (20011, 25, 64, 384, 'bf16'), (30001, 10, 128, 200, 'bf16'),
                                           (25000, 10, 64, 512, 'f32')])
def test_pool_kernel_many_tiles_per_cta(g, n, S, d, H, dtype):
    """linear_pool_ws_umma.cu over many tiles per CTA (ring wrap-around, accumulator-buffer parities, several staged id batches, a
    ragged last tile, an id outside the table), rows by id and rows in place; against fp64 on the same operands."""
    gen = torch.Generator().manual_seed(n + S + d)
    rows = 50000
    tdt = torch.bfloat16 if dtype == 'bf16' else torch.float32
    rnd = (lambda t: _bf16(t)) if dtype == 'bf16' else (lambda t: t)
    table = rnd(torch.randn((rows, d), generator=gen))
    w, b = rnd(torch.randn((H, d), generator=gen) / d ** 0.5), torch.randn((H,), generator=gen)
    ids = torch.randint(0, rows, (n * S,), generator=gen)
    ids[5] = rows + 7                                          # outside the table: reads as a zero row
    pad = lambda t: g.ops.pad_table(t.float(), tdt)[0][:, :t.shape[1]]
    rows_d = table[ids.clamp(max=rows - 1)].double()
    rows_d[5] = 0
    tol = 2e-4 if dtype == 'bf16' else 5e-3                    # fp32 operands run as TF32 products here
    for reduce in ('max', 'mean'):
        h = torch.relu(rows_d @ w.double().t() + b.double()).view(n, S, H)
        want = h.max(dim=1)[0] if reduce == 'max' else h.mean(dim=1)
        got = g.ops.linear_pooled(pad(table), pad(w), n, S, reduce, ids=ids.cuda(), bias=b.cuda())
        np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=tol, atol=tol)
    nb = rnd(torch.randn((n * S, d), generator=gen))
    h = torch.relu(nb.double() @ w.double().t() + b.double()).view(n, S, H)
    got = g.ops.linear_pooled(pad(nb), pad(w), n, S, 'max', bias=b.cuda())
    np.testing.assert_allclose(got.cpu().numpy(), h.max(dim=1)[0].numpy(), rtol=tol, atol=tol)


@pytest.mark.parametrize('n,d,O', [(777, 602, 128), (128 * 150 + 3, 64, 64), (50, 256, 128)])
def test_umma_concat_with_self(g, n, d, O):
    """[fc_x(table[ids]) | fc_neib(m)] + relu in one tensor-core launch (two TMEM accumulators per tile)."""
    gen = torch.Generator().manual_seed(n)
    table = _bf16(torch.randn((n + 99, d), generator=gen))
    m = _bf16(torch.randn((n, d), generator=gen))
    wx, wn = _bf16(torch.randn((O, d), generator=gen) / d ** 0.5), _bf16(torch.randn((O, d), generator=gen) / d ** 0.5)
    ids = torch.randint(0, n + 99, (n,), generator=gen)
    pad = lambda t: g.ops.pad_table(t.float(), torch.bfloat16)[0][:, :t.shape[1]]
    want = torch.relu(torch.cat([table[ids].double() @ wx.double().t(), m.double() @ wn.double().t()], dim=1))
    got = g.ops.linear([dict(a=pad(table), ids=ids.cuda(), w=pad(wx), col0=0), dict(a=pad(m), w=pad(wn), col0=O)], n,
                       act='relu', out_dtype=torch.float32, exact=False)
    assert got.shape == (n, 2 * O)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-4, atol=2e-4)


def test_umma_wide_output_is_split(g):
    """O = 512 (the pool MLP) does not fit one 256-column accumulator: the dispatcher issues column blocks."""
    n, d, O = 1500, 64, 512
    gen = torch.Generator().manual_seed(3)
    a, w, b = _bf16(torch.randn((n, d), generator=gen)), _bf16(torch.randn((O, d), generator=gen) / 8), torch.randn((O,), generator=gen)
    pad = lambda t: g.ops.pad_table(t.float(), torch.bfloat16)[0][:, :t.shape[1]]
    want = torch.relu(a.double() @ w.double().t() + b.double())
    got = g.ops.linear([dict(a=pad(a), w=pad(w), bias=b.cuda())], n, act='relu', exact=False)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize('n,d,O,S', [(300, 602, 128, 10), (129, 64, 64, 25), (1000, 256, 128, 3), (50, 100, 16, 40)])
def test_fused_gather_mean_projection(g, n, d, O, S):
    """[fc_x(table[ids]) | fc_neib(mean_j table[nb_ids])]: the gather+mean happens inside the projection kernel's
    operand load (tcgen05 path, bf16) -- compared with fp64 on the same bf16 operands, and with the fp32 FFMA path."""
    gen = torch.Generator().manual_seed(n + S)
    rows = 2000
    table = _bf16(torch.randn((rows, d), generator=gen))
    wx, wn = _bf16(torch.randn((O, d), generator=gen) / d ** 0.5), _bf16(torch.randn((O, d), generator=gen) / d ** 0.5)
    ids = torch.randint(0, rows, (n,), generator=gen)
    nb = torch.randint(0, rows, (n * S,), generator=gen)
    pad = lambda t: g.ops.pad_table(t.float(), torch.bfloat16)[0][:, :t.shape[1]]
    mean = table[nb].float().view(n, S, d).sum(dim=1) * (1.0 / S)               # fp32 sum in j order, like the kernel
    mean_bf16 = mean.to(torch.bfloat16)                                         # the kernel rounds the mean to bf16 in smem
    want = torch.relu(torch.cat([table[ids].double() @ wx.double().t(), mean_bf16.double() @ wn.double().t()], dim=1))
    segs = [dict(a=pad(table), ids=ids.cuda(), w=pad(wx), col0=0), dict(a=pad(table), ids=nb.cuda(), w=pad(wn), col0=O, S=S)]
    got = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact=False)
    # bf16 rounding of the mean can flip by one ulp vs the torch emulation (summation order): 2^-8 relative on that operand
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-3, atol=2e-3)
    exact = torch.relu(torch.cat([table[ids].double() @ wx.double().t(),
                                  table[nb].double().view(n, S, d).mean(dim=1) @ wn.double().t()], dim=1))
    got32 = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact=True)   # FFMA path: mean kept in fp32
    np.testing.assert_allclose(got32.cpu().numpy(), exact.numpy(), rtol=1e-4, atol=1e-5)


def test_fused_contiguous_mean_projection(g):
    """ids=None, S>1: the layer-2 case -- neighbours are rows r*S+j of the previous layer's output."""
    n, d, O, S = 200, 256, 128, 25
    gen = torch.Generator().manual_seed(8)
    h = _bf16(torch.randn((n + n * S, d), generator=gen))
    wx, wn = _bf16(torch.randn((O, d), generator=gen) / 16), _bf16(torch.randn((O, d), generator=gen) / 16)
    pad = lambda t: g.ops.pad_table(t.float(), torch.bfloat16)[0][:, :t.shape[1]]
    hd = pad(h)
    mean = (h[n:].float().view(n, S, d).sum(dim=1) * (1.0 / S)).to(torch.bfloat16)
    want = torch.cat([h[:n].double() @ wx.double().t(), mean.double() @ wn.double().t()], dim=1)
    got = g.ops.linear([dict(a=hd[:n], w=pad(wx), col0=0), dict(a=hd[n:], w=pad(wn), col0=O, S=S)], n, exact=False)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize('n,d,O', [(300, 64, 128), (1000, 256, 64), (129, 100, 16)])
def test_umma_tf32_projection(g, n, d, O):
    """fp32 operands on the tensor cores as TF32 (exact=False): products carry a 10-bit mantissa, fp32 accumulate.
    Stated tolerance: 2e-3 * sqrt(d) absolute on unit-variance operands (measured errors are ~10x smaller)."""
    gen = torch.Generator().manual_seed(d)
    table = torch.randn((n + 77, d), generator=gen)
    m = torch.randn((n, d), generator=gen)
    wx, wn = torch.randn((O, d), generator=gen) / d ** 0.5, torch.randn((O, d), generator=gen) / d ** 0.5
    ids = torch.randint(0, n + 77, (n,), generator=gen)
    pad = lambda t: g.ops.pad_table(t, torch.float32)[0][:, :t.shape[1]]
    segs = [dict(a=pad(table), ids=ids.cuda(), w=pad(wx), col0=0), dict(a=pad(m), w=pad(wn), col0=O)]
    want = torch.relu(torch.cat([table[ids].double() @ wx.double().t(), m.double() @ wn.double().t()], dim=1))
    got = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact=False)
    err = (got.cpu().double() - want).abs().max().item()
    assert err < 2e-3 * d ** 0.5, err
    assert err > 0 or d < 8                       # it really ran in reduced precision (the FFMA path would be ~1e-6)
    exact = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact=True)
    np.testing.assert_allclose(exact.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('n,H,h_dtype', [(1, 8, torch.float32), (37, 64, torch.float32), (1000, 512, torch.float32), (129, 64, torch.bfloat16)])
def test_lstm_cell_equals_torch(g, n, H, h_dtype):
    """gsage_lstm_cell against torch.nn.LSTMCell's arithmetic (gate order i, f, g, o), three chained steps from the zero state."""
    gen = torch.Generator().manual_seed(n + H)
    b_ih, b_hh = torch.randn((4 * H,), generator=gen), torch.randn((4 * H,), generator=gen)
    c_ref, h_ref = torch.zeros((n, H), dtype=torch.float64), torch.zeros((n, H), dtype=torch.float64)
    c = torch.full((n, H), 7.0, device='cuda')                          # garbage: the first step must not read it
    h = torch.full((n, H), 7.0, device='cuda').to(h_dtype)
    for t in range(3):
        gx, gh = torch.randn((n, 4 * H), generator=gen), torch.randn((n, 4 * H), generator=gen)
        gates = gx.double() + b_ih.double() + b_hh.double() + (gh.double() if t > 0 else 0.0)
        i, f, gg, o = gates[:, :H], gates[:, H:2 * H], gates[:, 2 * H:3 * H], gates[:, 3 * H:]
        c_ref = torch.sigmoid(f) * c_ref + torch.sigmoid(i) * torch.tanh(gg)
        h_ref = torch.sigmoid(o) * torch.tanh(c_ref)
        g.ops.lstm_cell(gx.cuda(), gh.cuda() if t > 0 else None, b_ih.cuda(), b_hh.cuda(), c, h, first=(t == 0))
        np.testing.assert_allclose(c.cpu().numpy(), c_ref.numpy(), rtol=1e-5, atol=1e-6)
        tol = dict(rtol=1e-5, atol=1e-6) if h_dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
        np.testing.assert_allclose(h.float().cpu().numpy(), h_ref.numpy(), **tol)


@pytest.mark.parametrize('n,S,d,H,gather', [(40, 10, 20, 64, True), (33, 25, 64, 32, False), (5, 3, 7, 8, True)])
def test_lstm_aggregator_equals_torch_lstm(g, n, S, d, H, gather):
    """operators.LSTMAggregator (narrow and id-taking entries) against the stock nn.LSTM it holds its parameters in
    (nn_modules.py:276-279: batch_first, last hidden state)."""
    torch.manual_seed(n + S)
    agg = g.aggregator_lookup['lstm'](input_dim=d, output_dim=16, activation=F.relu, hidden_dim=H)
    table = torch.randn((n * S + 50, d))
    x = torch.randn((n, d))
    with torch.no_grad():
        if gather:
            ids_self = torch.randint(0, table.shape[0], (n,))
            ids_nb = torch.randint(0, table.shape[0], (n * S,))
            xs, nb = table[ids_self], table[ids_nb]
        else:
            xs, nb = x, table[:n * S]
        seq, _ = agg.lstm(nb.view(n, S, d))
        want = F.relu(torch.cat([agg.fc_x(xs), agg.fc_neib(seq[:, -1])], dim=1))
    agg = agg.cuda()
    if gather:
        got = agg.forward_ids(table.cuda(), ids_self.cuda(), ids_nb.cuda(), S)
    else:
        got = agg(xs.cuda(), nb.cuda())
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('n,d,O', [(300, 64, 128), (128 * 200 + 5, 64, 128), (1000, 100, 64), (129, 256, 32), (777, 64, 16)])
def test_three_tf32_projection_has_fp32_accuracy(g, n, d, O, monkeypatch):
    """exact='x3': fp32 operands on the tensor cores as 3 x TF32 (a = a_hi + a_lo, w = w_hi + w_lo, three products, fp32
    accumulate).  Same bar as the FFMA kernel (rtol 1e-4 / atol 1e-5 against float64), and it must be a different kernel
    than both the FFMA one (GSAGE_FP32_FFMA=1) and the single-pass TF32 one (whose error is ~100x larger)."""
    gen = torch.Generator().manual_seed(n + d)
    table = torch.randn((n + 77, d), generator=gen)
    m = torch.randn((n, d), generator=gen)
    wx, wn = torch.randn((O, d), generator=gen) / d ** 0.5, torch.randn((O, d), generator=gen) / d ** 0.5
    bias = torch.randn((O,), generator=gen)
    ids = torch.randint(0, n + 77, (n,), generator=gen)
    pad = lambda t: g.ops.pad_table(t, torch.float32)[0][:, :t.shape[1]]
    segs = [dict(a=pad(table), ids=ids.cuda(), w=pad(wx), col0=0, bias=bias.cuda()), dict(a=pad(m), w=pad(wn), col0=O)]
    want = torch.relu(torch.cat([table[ids].double() @ wx.double().t() + bias.double(), m.double() @ wn.double().t()], dim=1))
    got = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact='x3').cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-4, atol=1e-5)
    err_x3 = (got.double() - want).abs().max().item()
    tf32 = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact=False).cpu()
    err_tf32 = (tf32.double() - want).abs().max().item()
    assert err_x3 < err_tf32 / 20, (err_x3, err_tf32)
    assert err_x3 < 4e-6 * d ** 0.5, err_x3
    monkeypatch.setenv('GSAGE_FP32_FFMA', '1')
    ffma = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact='x3').cpu()
    np.testing.assert_allclose(ffma.numpy(), want.numpy(), rtol=1e-4, atol=1e-5)
    assert not torch.equal(ffma, got)              # different summation order: the tensor-core path really ran


@pytest.mark.parametrize('n,S,d,H', [(100, 10, 64, 512), (37, 25, 64, 512), (513, 3, 100, 32), (8, 128, 64, 48), (3000, 10, 256, 512),
                                     (257, 25, 256, 512), (1000, 10, 602, 512), (50, 64, 64, 200), (64, 2, 72, 136)])
@pytest.mark.parametrize('reduce', ['max', 'mean'])
def test_umma_pooled_epilogue(g, n, S, d, H, reduce):
    """relu(MLP) on tcgen05 with the max / mean over the S neighbour rows of each parent done in the epilogue."""
    gen = torch.Generator().manual_seed(n * S + d)
    rows = 3000
    table = _bf16(torch.randn((rows, d), generator=gen))
    w, b = _bf16(torch.randn((H, d), generator=gen) / d ** 0.5), torch.randn((H,), generator=gen)
    ids = torch.randint(0, rows, (n * S,), generator=gen)
    pad = lambda t: g.ops.pad_table(t.float(), torch.bfloat16)[0][:, :t.shape[1]]
    h = torch.relu(table[ids].double() @ w.double().t() + b.double()).view(n, S, H)
    want = h.max(dim=1)[0] if reduce == 'max' else h.mean(dim=1)
    got = g.ops.linear_pooled(pad(table), pad(w), n, S, reduce, ids=ids.cuda(), bias=b.cuda())
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-4, atol=2e-4)
    # contiguous neighbours (ids=None): rows p*S+j of the operand itself
    nb = _bf16(torch.randn((n * S, d), generator=gen))
    h = torch.relu(nb.double() @ w.double().t() + b.double()).view(n, S, H)
    want = h.max(dim=1)[0] if reduce == 'max' else h.mean(dim=1)
    got = g.ops.linear_pooled(pad(nb), pad(w), n, S, reduce, bias=b.cuda(), out_dtype=torch.bfloat16)
    np.testing.assert_allclose(got.float().cpu().numpy(), want.numpy(), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('n,S,d,H,dtype', [(40003, 10, 64, 512, 'bf16'), (20011, 25, 64, 384, 'bf16'), (30001, 10, 128, 200, 'bf16'),
                                           (25000, 10, 64, 512, 'f32')])
@pytest.mark.parametrize('fill', ['lsu', 'tma'])
def test_pool_kernel_128_column_tiles(g, n, S, d, H, dtype, fill, monkeypatch):
    """linear_pool_n128_umma.cu over many tiles per CTA (ring wrap-around, accumulator-slot parities, several id batches, a ragged
    last tile), rows by id through either fill path and rows in place; against fp64 on the same operands, and bit-for-bit against
    the 64-column kernel (same MMAs in the same k order)."""
    monkeypatch.setenv('GSAGE_POOL_FILL', fill)
    gen = torch.Generator().manual_seed(n + S + d)
    rows = 50000
    tdt = torch.bfloat16 if dtype == 'bf16' else torch.float32
    rnd = (lambda t: _bf16(t)) if dtype == 'bf16' else (lambda t: t)
    table = rnd(torch.randn((rows, d), generator=gen))
    w, b = rnd(torch.randn((H, d), generator=gen) / d ** 0.5), torch.randn((H,), generator=gen)
    ids = torch.randint(0, rows, (n * S,), generator=gen)
    ids[5] = rows + 7                                          # outside the table: reads as a zero row
    pad = lambda t: g.ops.pad_table(t.float(), tdt)[0][:, :t.shape[1]]
    rows_d = table[ids.clamp(max=rows - 1)].double()
    rows_d[5] = 0
    tol = 2e-4 if dtype == 'bf16' else 5e-3                    # fp32 operands run as TF32 products here
    for reduce in ('max', 'mean'):
        h = torch.relu(rows_d @ w.double().t() + b.double()).view(n, S, H)
        want = h.max(dim=1)[0] if reduce == 'max' else h.mean(dim=1)
        got = g.ops.linear_pooled(pad(table), pad(w), n, S, reduce, ids=ids.cuda(), bias=b.cuda())
        np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=tol, atol=tol)
        monkeypatch.setenv('GSAGE_NO_POOL_N128', '1')
        ref = g.ops.linear_pooled(pad(table), pad(w), n, S, reduce, ids=ids.cuda(), bias=b.cuda())
        monkeypatch.delenv('GSAGE_NO_POOL_N128')
        if reduce == 'max':
            assert torch.equal(got, ref)                       # (the mean sums the S rows in the same order too, but fma contraction may differ)
        else:
            np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=1e-6, atol=1e-6)
    nb = rnd(torch.randn((n * S, d), generator=gen))
    h = torch.relu(nb.double() @ w.double().t() + b.double()).view(n, S, H)
    got = g.ops.linear_pooled(pad(nb), pad(w), n, S, 'max', bias=b.cuda())
    np.testing.assert_allclose(got.cpu().numpy(), h.max(dim=1)[0].numpy(), rtol=tol, atol=tol)


@pytest.mark.parametrize('n,d,O', [(777, 602, 128), (128 * 150 + 3, 64, 64), (50, 256, 128), (128 * 300 + 1, 602, 128), (3000, 256, 256)])
def test_weight_stationary_kernel_equals_streaming_kernel(g, n, d, O, monkeypatch):
    """linear_ws_umma.cu (W of a phase resident in shared memory; d=602 runs as two phases, one segment each) against
    linear_umma.cu (W streamed with every stage): same MMAs in the same k order -> bit-identical fp32 results, and
    both within 2e-4 of fp64 on the bf16 operands."""
    gen = torch.Generator().manual_seed(n + d)
    table = _bf16(torch.randn((n + 99, d), generator=gen))
    m = _bf16(torch.randn((n, d), generator=gen))
    wx, wn = _bf16(torch.randn((O, d), generator=gen) / d ** 0.5), _bf16(torch.randn((O, d), generator=gen) / d ** 0.5)
    bx = torch.randn((O,), generator=gen)
    ids = torch.randint(0, n + 99, (n,), generator=gen)
    pad = lambda t: g.ops.pad_table(t.float(), torch.bfloat16)[0][:, :t.shape[1]]
    segs = [dict(a=pad(table), ids=ids.cuda(), w=pad(wx), col0=0, bias=bx.cuda()), dict(a=pad(m), w=pad(wn), col0=O)]
    want = torch.relu(torch.cat([table[ids].double() @ wx.double().t() + bx.double(), m.double() @ wn.double().t()], dim=1))
    got_ws = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact=False)
    monkeypatch.setenv('GSAGE_NO_WS', '1')
    got_stream = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact=False)
    monkeypatch.delenv('GSAGE_NO_WS')
    np.testing.assert_allclose(got_ws.cpu().numpy(), want.numpy(), rtol=2e-4, atol=2e-4)
    if 2 * O <= 256:                       # wider pairs are split differently by the streaming dispatcher
        np.testing.assert_array_equal(got_ws.cpu().numpy(), got_stream.cpu().numpy())
    out16 = g.ops.linear(segs, n, act='relu', out_dtype=torch.bfloat16, exact=False)
    np.testing.assert_allclose(out16.float().cpu().numpy(), want.numpy(), rtol=1e-2, atol=1e-2)


# ---- fused attention reduction (attention_umma.cu) -----------------------------------------------------------------
# Tolerance: bf16 operands are exact; scores go through a bf16 tensor-core product (fp32 accumulate), tanh.approx.f32
# (~5e-4 relative) and __expf; the weighted sum is fp32 over bf16 rows and is rounded to bf16 on store.  Against fp64 on
# the same bf16 operands: rtol 2e-2 / atol 2e-2 on unit-variance rows (measured ~3e-3).

@pytest.mark.parametrize('n,S,d,gather', [(300, 10, 256, True), (77, 25, 256, False), (1000, 10, 602, True), (64, 2, 64, True),
                                          (5, 128, 100, True), (513, 3, 8, False), (20000, 10, 256, True), (64, 8, 64, True), (5, 32, 100, True)])
def test_fused_attention_aggregate(g, n, S, d, gather):
    gen = torch.Generator().manual_seed(n + S + d)
    H = 32
    rows = 3000 if gather else n * S
    table = _bf16(torch.randn((rows, d), generator=gen))
    table[0] = 0                                                  # the dummy node: a zero row that is NOT masked
    w1 = _bf16(torch.randn((H, d), generator=gen) / d ** 0.5)
    b1 = torch.randn((H,), generator=gen) * 0.1
    w2 = torch.randn((H, H), generator=gen) / H ** 0.5
    xa = torch.randn((n, H), generator=gen)
    ids = torch.randint(0, rows, (n * S,), generator=gen) if gather else None
    if gather:
        ids[::7] = 0
    nb = (table[ids] if gather else table).double().view(n, S, d)
    a_n = torch.tanh(nb @ w1.double().t() + b1.double()) @ w2.double().t()          # (n, S, H)
    sc = torch.softmax((a_n * xa.double().view(n, 1, H)).sum(-1), dim=1)            # (n, S)
    want = (sc.unsqueeze(-1) * nb).sum(1)
    pad = lambda t: g.ops.pad_table(t.float(), torch.bfloat16)[0][:, :t.shape[1]]
    for out_dtype in (torch.float32, torch.bfloat16):
        got = g.ops.attention_aggregate(pad(table), None if ids is None else ids.cuda(), n, S, pad(w1), w2.cuda(), xa.cuda(), b1=b1.cuda(),
                                        out_dtype=out_dtype)
        np.testing.assert_allclose(got.float().cpu().numpy(), want.numpy(), rtol=2e-2, atol=2e-2)
    # no bias (the plain reference module has none)
    a_n = torch.tanh(nb @ w1.double().t()) @ w2.double().t()
    sc = torch.softmax((a_n * xa.double().view(n, 1, H)).sum(-1), dim=1)
    want = (sc.unsqueeze(-1) * nb).sum(1)
    got = g.ops.attention_aggregate(pad(table), None if ids is None else ids.cuda(), n, S, pad(w1), w2.cuda(), xa.cuda(), out_dtype=torch.float32)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-2, atol=2e-2)


def test_fused_attention_rejects_fp32_tables(g):
    with pytest.raises(ValueError):
        g.ops.attention_aggregate(torch.randn(100, 64).cuda(), None, 10, 10, torch.randn(32, 64).cuda(), torch.randn(32, 32).cuda(),
                                  torch.randn(10, 32).cuda())


def test_tensor_core_gathers_read_out_of_range_ids_as_zero_rows(g):
    """ids outside the table are zero rows in the TMA gathers too (the tensor maps are bounded by the table's rows):
    same behaviour as gather_reduce.cu, and no access past the table's last row."""
    gen = torch.Generator().manual_seed(1)
    rows, d, O, n = 37, 64, 128, 300
    table = _bf16(torch.randn((rows, d), generator=gen))
    w = _bf16(torch.randn((O, d), generator=gen) / 8)
    b = torch.randn((O,), generator=gen)
    ids = torch.randint(0, rows, (n,), generator=gen)
    ids[::5] = rows + 1000                                    # far past the table
    ids[3] = -7                                               # negative: also outside
    pad = lambda t: g.ops.pad_table(t.float(), torch.bfloat16)[0][:, :t.shape[1]]
    src = table.double()[ids.clamp(0, rows - 1)]
    src[(ids < 0) | (ids >= rows)] = 0
    want = src @ w.double().t() + b.double()
    got = g.ops.linear([dict(a=pad(table), ids=ids.cuda(), w=pad(w), bias=b.cuda())], n, exact=False)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-4, atol=2e-4)
