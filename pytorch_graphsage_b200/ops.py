"""
Tensor-level wrappers over the C ABI (include/gsage_b200.h).  torch is plumbing only: it owns the device
memory and the stream; every arithmetic step happens inside libgsage_b200.so.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import F32, BF16, check, lib

_bound_device = [None]


def _bind_device(t=None):
    """One process per GPU: point the library's CUDA runtime at torch's current device."""
    if not torch.cuda.is_available():
        raise _lib.GsageError('pytorch_graphsage_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
    dev = t.device.index if (t is not None and t.is_cuda and t.device.index is not None) else torch.cuda.current_device()
    if _bound_device[0] != dev:
        check(lib().gsage_set_device(dev))
        _bound_device[0] = dev
    return dev


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError('gsage: only float32 / bfloat16 tensors are supported, got %s' % t.dtype)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _rows2d(t):
    assert t.dim() == 2 and t.stride(1) == 1, 'gsage: expected a row-major 2-D tensor'
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def pad_table(t, dtype=None):
    """(rows, d) -> device tensor whose rows start 16-byte aligned with zero padding (a view of width d is returned
    together with the leading dimension).  Returns (storage (rows, ld), d)."""
    t = torch.as_tensor(t)
    dtype = dtype or (t.dtype if t.dtype in (torch.float32, torch.bfloat16) else torch.float32)
    per = 16 // (2 if dtype == torch.bfloat16 else 4)
    d = t.shape[1]
    ld = (d + 2 * per - 1) // (2 * per) * (2 * per)            # 32-byte rows: whole DRAM sectors
    out = torch.zeros((t.shape[0], ld), dtype=dtype, device='cuda')
    out[:, :d] = t.to(device='cuda', dtype=dtype)
    return out, d


def aligned_rows(t):
    """The library streams rows with 16-byte loads: rows must start 16-byte aligned and be zero-padded to a whole
    chunk.  Tensors that already qualify pass through; anything else gets one padded copy."""
    es = t.element_size()
    per = 16 // es
    if t.dim() == 2 and t.stride(1) == 1 and t.data_ptr() % 16 == 0 and (t.stride(0) * es) % 16 == 0 \
            and (t.shape[1] % per == 0 or t.stride(0) >= (t.shape[1] + per - 1) // per * per):
        return t
    store, d = pad_table(t, t.dtype)
    return store[:, :d]


def _ids_arg(ids):
    if ids is None:
        return None
    assert ids.is_cuda and ids.dtype == torch.int64 and ids.is_contiguous(), 'gsage: ids must be a contiguous CUDA int64 tensor'
    return ptr(ids)


def gather_reduce(table, ids, n_parents, S, reduce='mean', weights=None, d=None, out_dtype=None, out=None):
    """out[p] = reduce_j w[p*S+j] * table[ids[p*S+j]]  (ids None: rows p*S+j of `table`)."""
    _bind_device(table)
    d = table.shape[1] if d is None else d
    table = aligned_rows(table)
    out_dtype = out_dtype or table.dtype
    if out is None:
        per = 16 // (2 if out_dtype == torch.bfloat16 else 4)
        ld_out = (d + per - 1) // per * per
        store = torch.zeros((n_parents, ld_out), dtype=out_dtype, device=table.device)
        out = store[:, :d]
    if weights is not None:
        assert weights.is_cuda and weights.dtype == torch.float32 and weights.is_contiguous()
    check(lib().gsage_gather_reduce(ptr(table), dt(table), _rows2d(table), table.shape[0], d, _ids_arg(ids), n_parents, S,
                                    _lib.REDUCE[reduce], ptr(weights), ptr(out), dt(out), _rows2d(out), stream()))
    return out


def gather_rows(table, ids, d=None, out_dtype=None):
    return gather_reduce(table, ids, ids.shape[0], 1, 'sum', d=d, out_dtype=out_dtype)


def linear(segments, n, act=None, out=None, out_dtype=torch.float32, exact=True):
    """segments: list of dicts(a=, w=, ids=None, bias=None, col0=0, d=None, S=1, w_trans=False).  One launch, <= 2 column ranges.
    S > 1: row r of that segment's operand is mean_j a[ids[r*S + j]] (fused gather+mean).
    w_trans: out = a . w instead of a . w^T (w is (d, O): the data gradient of a Linear; fp32 FFMA kernel only)."""
    segs = (_lib.LinearSeg * len(segments))()
    width = 0
    for i, sgm in enumerate(segments):
        a, w = sgm['a'], sgm['w']
        _bind_device(a)
        d = sgm.get('d') or (w.shape[0] if sgm.get('w_trans') else w.shape[1])
        assert w.is_cuda and w.dim() == 2 and w.stride(1) == 1
        bias = sgm.get('bias')
        segs[i] = _lib.LinearSeg(ptr(a), dt(a), _rows2d(a), _ids_arg(sgm.get('ids')), ptr(w), dt(w), _rows2d(w), d,
                                 (w.shape[1] if sgm.get('w_trans') else w.shape[0]), ptr(bias), sgm.get('col0', 0), int(sgm.get('S', 1)),
                                 1 if sgm.get('w_trans') else 0, a.shape[0])
        width = max(width, sgm.get('col0', 0) + (w.shape[1] if sgm.get('w_trans') else w.shape[0]))
    if out is None:
        out = torch.empty((n, width), dtype=out_dtype, device=segments[0]['a'].device)
    check(lib().gsage_linear(segs, len(segments), n, _lib.ACT[act], ptr(out), dt(out), _rows2d(out), 2 if exact == 'x3' else (1 if exact else 0), stream()))
    return out


def linear_pooled(a, w, n_parents, S, reduce='max', ids=None, bias=None, act='relu', out_dtype=torch.float32):
    """reduce_j act(a[ids[p*S+j]] . w^T + bias): the pool aggregators' per-neighbour MLP with the pool fused in the epilogue."""
    _bind_device(a)
    seg = _lib.LinearSeg(ptr(a), dt(a), _rows2d(a), _ids_arg(ids), ptr(w), dt(w), _rows2d(w), w.shape[1], w.shape[0], ptr(bias), 0, 1, 0,
                         a.shape[0])
    out = torch.empty((n_parents, w.shape[0]), dtype=out_dtype, device=a.device)
    check(lib().gsage_linear_pooled(C.byref(seg), n_parents, S, _lib.REDUCE[reduce], _lib.ACT[act], ptr(out), dt(out), _rows2d(out), stream()))
    return out


def wgrad(g, a, ids=None, n=None, exact=True):
    """dW (O, d) fp32 = g[:n]^T . a[ids] (ids None: a[:n]) -- the weight gradient of a Linear (gsage_wgrad)."""
    _bind_device(a)
    n = g.shape[0] if n is None else n
    O, d = g.shape[1], a.shape[1]
    dw = torch.empty((O, d), dtype=torch.float32, device=a.device)
    check(lib().gsage_wgrad(ptr(g), dt(g), _rows2d(g), O, ptr(a), dt(a), _rows2d(a), a.shape[0], _ids_arg(ids), d, n, ptr(dw), d,
                            1 if exact else 0, stream()))
    return dw


# ---- gradient kernels of the narrow plug-in API (narrow_backward.cu); everything fp32 ------------------------------
def _f32c(t):
    assert t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1, 'gsage: expected a row-major fp32 CUDA matrix'
    return t


def act_backward(dout, out, act):
    """dout * act'(out) for act in (None, 'relu', 'tanh'); `out` is the post-activation value."""
    if act is None:
        return dout
    dout, out = _f32c(dout), _f32c(out)
    _bind_device(dout)
    dpre = torch.empty(dout.shape, dtype=torch.float32, device=dout.device)
    check(lib().gsage_act_backward(ptr(dout), _rows2d(dout), ptr(out), _rows2d(out), dout.shape[0], dout.shape[1], _lib.ACT[act],
                                   ptr(dpre), _rows2d(dpre), stream()))
    return dpre


def segment_broadcast(src, S, scale=1.0):
    src = _f32c(src)
    _bind_device(src)
    dst = torch.empty((src.shape[0] * S, src.shape[1]), dtype=torch.float32, device=src.device)
    check(lib().gsage_segment_broadcast(ptr(src), _rows2d(src), src.shape[0], src.shape[1], S, float(scale), ptr(dst), _rows2d(dst), stream()))
    return dst


def segment_max_backward(h, dpooled, S):
    h, dpooled = _f32c(h), _f32c(dpooled)
    _bind_device(h)
    dh = torch.empty(h.shape, dtype=torch.float32, device=h.device)
    check(lib().gsage_segment_max_backward(ptr(h), _rows2d(h), ptr(dpooled), _rows2d(dpooled), dpooled.shape[0], S, h.shape[1], ptr(dh), _rows2d(dh),
                                           stream()))
    return dh


def attention_sum_backward(neibs, dm, w, na, xa, S):
    """-> (dneibs: the direct path w_j dM_p, dna, dxa)."""
    neibs, dm, na, xa = _f32c(neibs), _f32c(dm), _f32c(na.contiguous()), _f32c(xa.contiguous())
    _bind_device(neibs)
    n, d, H = dm.shape[0], neibs.shape[1], na.shape[1]
    dn = torch.empty((n * S, d), dtype=torch.float32, device=neibs.device)
    dna, dxa = torch.empty_like(na), torch.empty_like(xa)
    scratch = torch.empty((n * S,), dtype=torch.float32, device=neibs.device)
    check(lib().gsage_attention_sum_backward(ptr(neibs), _rows2d(neibs), d, n, S, ptr(dm), _rows2d(dm), ptr(w), ptr(na), ptr(xa), H, ptr(dn), _rows2d(dn),
                                             ptr(dna), ptr(dxa), ptr(scratch), stream()))
    return dn, dna, dxa


def colsum(x):
    x = _f32c(x.contiguous())
    _bind_device(x)
    out = torch.empty((x.shape[1],), dtype=torch.float32, device=x.device)
    check(lib().gsage_colsum(ptr(x), x.shape[0], x.shape[1], ptr(out), stream()))
    return out


def embedding_backward(drows, ids, table_rows):
    drows = _f32c(drows)
    _bind_device(drows)
    grad = torch.zeros((table_rows, drows.shape[1]), dtype=torch.float32, device=drows.device)
    check(lib().gsage_embedding_backward(ptr(drows), _rows2d(drows), drows.shape[1], _ids_arg(ids), ids.shape[0], ptr(grad), grad.shape[1], table_rows,
                                         stream()))
    return grad


def l2_normalize_backward(z, dzn):
    z, dzn = _f32c(z.contiguous()), _f32c(dzn.contiguous())
    _bind_device(z)
    dz = torch.empty_like(z)
    check(lib().gsage_l2_normalize_backward(ptr(z), ptr(dzn), z.shape[0], z.shape[1], ptr(dz), stream()))
    return dz


def attention_weights(na, xa, n_parents, S):
    _bind_device(na)
    w = torch.empty((n_parents * S,), dtype=torch.float32, device=na.device)
    check(lib().gsage_attention_weights(ptr(na), ptr(xa), dt(na), _rows2d(na), na.shape[1], n_parents, S, ptr(w), stream()))
    return w


def gather_mean_project(table, ids, n_parents, S, w, bias=None, act=None, out=None, col0=0, out_dtype=torch.bfloat16):
    """act(mean_j table[ids[p*S+j]] . w^T + bias) in one kernel (gsage_gather_mean_project)."""
    _bind_device(table)
    if out is None:
        out = torch.empty((n_parents, col0 + w.shape[0]), dtype=out_dtype, device=table.device)
    check(lib().gsage_gather_mean_project(ptr(table), dt(table), _rows2d(table), table.shape[0], table.shape[1], _ids_arg(ids), n_parents, S,
                                          ptr(w), dt(w), _rows2d(w), w.shape[0], ptr(bias), _lib.ACT[act], ptr(out), dt(out), _rows2d(out),
                                          col0, stream()))
    return out


def lstm_cell(gx, gh, b_ih, b_hh, c, h, first):
    """One LSTM time step in place (gsage_lstm_cell): gates = gx + gh + b_ih + b_hh (n, 4H) fp32, torch gate order i, f, g, o;
    c (n, H) fp32 and h (n, H) fp32 | bf16 are updated.  first=True: zero initial state (c not read, gh ignored).
    gx may be a strided view (step t of an (n, S, 4H) block)."""
    _bind_device(gx)
    n, H = c.shape
    assert gx.dtype == torch.float32 and c.dtype == torch.float32 and gx.shape[1] == 4 * H
    check(lib().gsage_lstm_cell(ptr(gx), _rows2d(gx), ptr(gh) if gh is not None else None, _rows2d(gh) if gh is not None else 0,
                                ptr(b_ih), ptr(b_hh), ptr(c), ptr(h), dt(h), _rows2d(h), n, H, 1 if first else 0, stream()))
    return h


def lstm_cell_backward(gx, gh, b_ih, b_hh, c_prev, dh, dc, dgates, first):
    """One step of BPTT through the cell (gsage_lstm_cell_backward): dc is updated in place to d loss / d c_{t-1}, dgates
    (n, 4H; may be a strided view) receives d loss / d pre-activation gates."""
    _bind_device(gx)
    n, H = dh.shape
    check(lib().gsage_lstm_cell_backward(ptr(gx), _rows2d(gx), ptr(gh) if gh is not None else None, _rows2d(gh) if gh is not None else 0,
                                         ptr(b_ih), ptr(b_hh), ptr(c_prev) if c_prev is not None else None, ptr(dh), ptr(dc), ptr(dgates),
                                         _rows2d(dgates), n, H, 1 if first else 0, stream()))
    return dgates


def attention_aggregate(table, ids, n_parents, S, w1, w2, xa, b1=None, out_dtype=None):
    """sum_j softmax_j(<a(n_pj), xa[p]>) n_pj in one launch (gsage_attention_aggregate); returns (n_parents, d)."""
    _bind_device(table)
    d = table.shape[1]
    out_dtype = out_dtype or table.dtype
    per = 32 // (2 if out_dtype == torch.bfloat16 else 4)          # rows of whole 32-byte sectors, zero padding (as pad_table lays them out)
    out = torch.zeros((n_parents, (d + per - 1) // per * per), dtype=out_dtype, device=table.device)[:, :d]
    check(lib().gsage_attention_aggregate(ptr(table), dt(table), _rows2d(table), table.shape[0], d, _ids_arg(ids), n_parents, S, ptr(w1), dt(w1), _rows2d(w1),
                                          w1.shape[0], ptr(b1), ptr(w2), ptr(xa), ptr(out), dt(out), _rows2d(out), stream()))
    return out


def l2_normalize(x):
    _bind_device(x)
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    check(lib().gsage_l2_normalize(ptr(x), dt(x), _rows2d(x), x.shape[0], x.shape[1], ptr(out), _rows2d(out), stream()))
    return out


def device_info():
    _bind_device()
    name = C.create_string_buffer(128)
    sms, cc1, cc2, mem = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
    check(lib().gsage_device_info(name, 128, C.byref(sms), C.byref(mem), C.byref(cc1), C.byref(cc2)))
    return dict(name=name.value.decode(), sm_count=sms.value, hbm_bytes=mem.value, cc=(cc1.value, cc2.value))


def u32_to_numpy(t):
    """int32-typed device tensor holding uint32 bit patterns -> numpy uint32."""
    return t.cpu().numpy().view(np.uint32)


class _DevView(object):
    def __init__(self, ptr_value, n, typestr):
        self.__cuda_array_interface__ = dict(shape=(n,), typestr=typestr, data=(ptr_value, False), version=2, strides=None)


def device_view(ptr_value, n, dtype_code):
    """Zero-copy torch view of `n` elements of library-owned device memory (dtype_code: -1 int64, 0 f32, 1 bf16)."""
    if dtype_code == -1:
        return torch.as_tensor(_DevView(ptr_value, n, '<i8'), device='cuda')
    if dtype_code == F32:
        return torch.as_tensor(_DevView(ptr_value, n, '<f4'), device='cuda')
    return torch.as_tensor(_DevView(ptr_value, n, '<i2'), device='cuda').view(torch.bfloat16)
