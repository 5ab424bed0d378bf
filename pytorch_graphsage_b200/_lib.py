"""
ctypes binding of libgsage_b200.so (include/gsage_b200.h).  There is NO fallback: if the library is missing
or a call fails, this raises -- the product path never routes through a CPU implementation.
"""

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# GSAGE_B200_LIB: load another build of the same library (profiles/scripts/build_timing_lib.sh: in-kernel cycle counters)
LIB_PATH = os.environ.get('GSAGE_B200_LIB') or os.path.join(HERE, 'libgsage_b200.so')

F32, BF16 = 0, 1
ABI_VERSION = 3
PROF_CATS = ('forward', 'sample', 'reduce', 'project', 'app0', 'layer2', 'head', 'wait')
ACT = {'none': 0, None: 0, 'identity': 0, 'relu': 1, 'tanh': 2}
REDUCE = {'mean': 0, 'max': 1, 'sum': 2}
AGGREGATOR = {'mean': 0, 'max_pool': 1, 'mean_pool': 2, 'attention': 3, 'lstm': 4}
PREP = {'identity': 0, 'node_embedding': 1, 'linear': 2}

ERR_INVALID, ERR_CUDA, ERR_INDEX, ERR_RNG, ERR_NOMEM = -1, -2, -3, -4, -5

c_i64 = C.c_int64
c_p = C.c_void_p


class LinearSeg(C.Structure):
    _fields_ = [('a_dev', c_p), ('a_dtype', C.c_int), ('lda', c_i64), ('ids_dev', c_p),
                ('w_dev', c_p), ('w_dtype', C.c_int), ('ldw', c_i64), ('d', C.c_int), ('O', C.c_int),
                ('bias_dev', c_p), ('col0', c_i64), ('reduce_S', C.c_int), ('w_transposed', C.c_int), ('a_rows', c_i64)]


class EngineConfig(C.Structure):
    _fields_ = [('aggregator', C.c_int), ('prep', C.c_int), ('n_layers', C.c_int),
                ('fanout', C.c_int * 2), ('out_dim', C.c_int * 2), ('act', C.c_int * 2),
                ('n_classes', C.c_int), ('compute_dtype', C.c_int),
                ('feats_dev', c_p), ('feats_dtype', C.c_int), ('feats_ld', c_i64), ('feats_dim', C.c_int),
                ('feats_rows', c_i64),
                ('emb_dev', c_p), ('emb_dtype', C.c_int), ('emb_ld', c_i64), ('emb_dim', C.c_int), ('n_nodes', c_i64),
                ('hidden_dim', C.c_int), ('max_batch', c_i64), ('allow_tf32', C.c_int)]


class LayerWeights(C.Structure):
    _fields_ = [('fc_x', c_p), ('fc_neib', c_p), ('mlp_w', c_p), ('mlp_b', c_p), ('att_w1', c_p), ('att_w2', c_p),
                ('lstm_w_ih', c_p), ('lstm_w_hh', c_p), ('lstm_b_ih', c_p), ('lstm_b_hh', c_p)]


class Grads(C.Structure):
    _fields_ = [('fc_x', c_p * 2), ('fc_neib', c_p * 2), ('fc_w', c_p), ('fc_b', c_p)]


class PoolGrads(C.Structure):
    _fields_ = [('mlp_w', c_p * 2), ('mlp_b', c_p * 2)]


class AttentionGrads(C.Structure):
    _fields_ = [('att_w1', c_p * 2), ('att_w2', c_p * 2)]


class PoolEmbeddingGrads(C.Structure):
    _fields_ = [('csum_x', c_p), ('d_table', c_p)]


class EmbeddingGrads(C.Structure):
    _fields_ = [('gx_raw', c_p), ('gn_raw', c_p), ('csum', c_p), ('d_table', c_p)]


class LinearPrepGrads(C.Structure):
    _fields_ = [('gx_raw', c_p), ('gn_raw', c_p)]


class Weights(C.Structure):
    _fields_ = [('layer', LayerWeights * 2), ('fc_w', c_p), ('fc_b', c_p), ('prep_fc_w', c_p), ('prep_fc_b', c_p),
                ('prep_out_dim', C.c_int)]


_SIGNATURES = {
    'gsage_abi_version': (C.c_int, []),
    'gsage_last_error': (C.c_char_p, []),
    'gsage_device_info': (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(c_i64), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'gsage_set_device': (C.c_int, [C.c_int]),
    'gsage_launch_count': (c_i64, []),
    'gsage_graph_from_csr': (C.c_int, [c_p, c_p, c_p, c_i64, c_i64, C.POINTER(c_p)]),
    'gsage_graph_from_triplets': (C.c_int, [c_p, c_p, c_p, c_i64, C.POINTER(c_p)]),
    'gsage_graph_destroy': (None, [c_p]),
    'gsage_graph_info': (C.c_int, [c_p, C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(C.c_int), C.POINTER(c_i64)]),
    'gsage_graph_degrees_host': (C.c_int, [c_p, c_p]),
    'gsage_graph_check': (C.c_int, [c_p, c_p]),
    'gsage_rng_create': (C.c_int, [C.POINTER(c_p)]),
    'gsage_rng_destroy': (None, [c_p]),
    'gsage_rng_seed': (C.c_int, [c_p, C.c_uint32, c_p]),
    'gsage_rng_set_state': (C.c_int, [c_p, c_p, C.c_int, c_p]),
    'gsage_rng_get_state': (C.c_int, [c_p, c_p, C.POINTER(C.c_int), c_p]),
    'gsage_rng_raw': (C.c_int, [c_p, c_i64, c_p, c_p]),
    'gsage_rng_randint': (C.c_int, [c_p, C.c_uint32, c_i64, c_p, c_p]),
    'gsage_rng_permutation': (C.c_int, [c_p, c_i64, c_p, c_p]),
    'gsage_rng_check': (C.c_int, [c_p, c_p]),
    'gsage_rng_consumed': (C.c_int, [c_p, C.POINTER(c_i64), c_p]),
    'gsage_sample_sparse': (C.c_int, [c_p, c_p, c_i64, C.c_int, c_p, c_p, c_p]),
    'gsage_sample_sparse_rng': (C.c_int, [c_p, c_p, c_p, c_i64, C.c_int, c_p, c_p]),
    'gsage_sample_dense': (C.c_int, [c_p, c_i64, C.c_int, c_p, c_i64, c_p, C.c_int, c_p, c_p]),
    'gsage_gather_rows': (C.c_int, [c_p, C.c_int, c_i64, c_i64, C.c_int, c_p, c_i64, c_p, C.c_int, c_i64, c_p]),
    'gsage_gather_reduce': (C.c_int, [c_p, C.c_int, c_i64, c_i64, C.c_int, c_p, c_i64, C.c_int, C.c_int, c_p, c_p, C.c_int, c_i64, c_p]),
    'gsage_attention_weights': (C.c_int, [c_p, c_p, C.c_int, c_i64, C.c_int, c_i64, C.c_int, c_p, c_p]),
    'gsage_attention_aggregate': (C.c_int, [c_p, C.c_int, c_i64, c_i64, C.c_int, c_p, c_i64, C.c_int, c_p, C.c_int, c_i64, C.c_int, c_p, c_p, c_p, c_p,
                                            C.c_int, c_i64, c_p]),
    'gsage_gather_mean_project': (C.c_int, [c_p, C.c_int, c_i64, c_i64, C.c_int, c_p, c_i64, C.c_int, c_p, C.c_int, c_i64, C.c_int, c_p, C.c_int, c_p,
                                            C.c_int, c_i64, c_i64, c_p]),
    'gsage_lstm_cell': (C.c_int, [c_p, c_i64, c_p, c_i64, c_p, c_p, c_p, c_p, C.c_int, c_i64, c_i64, C.c_int, C.c_int, c_p]),
    'gsage_lstm_cell_backward': (C.c_int, [c_p, c_i64, c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_i64, C.c_int, C.c_int, c_p]),
    'gsage_l2_normalize': (C.c_int, [c_p, C.c_int, c_i64, c_i64, C.c_int, c_p, c_i64, c_p]),
    'gsage_linear': (C.c_int, [C.POINTER(LinearSeg), C.c_int, c_i64, C.c_int, c_p, C.c_int, c_i64, C.c_int, c_p]),
    'gsage_linear_pooled': (C.c_int, [C.POINTER(LinearSeg), c_i64, C.c_int, C.c_int, C.c_int, c_p, C.c_int, c_i64, c_p]),
    'gsage_engine_create': (C.c_int, [C.POINTER(EngineConfig), C.POINTER(c_p)]),
    'gsage_engine_destroy': (None, [c_p]),
    'gsage_engine_set_weights': (C.c_int, [c_p, C.POINTER(Weights), c_p]),
    'gsage_engine_forward': (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_p, c_p]),
    'gsage_engine_forward_sharded': (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_p, c_p]),
    'gsage_engine_forward_host': (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_p, c_p]),
    'gsage_engine_forward_dense': (C.c_int, [c_p, c_p, c_i64, C.c_int, c_p, c_p, c_p, c_i64, c_p, c_p]),
    'gsage_engine_forward_host_next': (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_p, c_i64, c_p, c_p]),
    'gsage_engine_sample_ahead': (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_p]),
    'gsage_engine_sample_ahead_host': (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_p]),
    'gsage_engine_sample_ahead_pending': (C.c_int, [c_p]),
    'gsage_engine_inputs_ready': (C.c_int, [c_p, c_p]),
    'gsage_engine_poll_errors': (C.c_int, [c_p]),
    'gsage_engine_peek': (C.c_int, [c_p, C.c_int, C.POINTER(c_p), C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(C.c_int)]),
    'gsage_engine_workspace_bytes': (c_i64, [c_p]),
    'gsage_engine_backward_head': (C.c_int, [c_p, c_p, C.POINTER(Grads), c_p]),
    'gsage_engine_backward_layer1': (C.c_int, [c_p, C.POINTER(Grads), c_p]),
    'gsage_engine_backward_pool': (C.c_int, [c_p, c_p, C.POINTER(Grads), C.POINTER(PoolGrads), c_p]),
    'gsage_engine_backward_attention': (C.c_int, [c_p, c_p, C.POINTER(Grads), C.POINTER(AttentionGrads), c_p]),
    'gsage_engine_backward_pool_embedding': (C.c_int, [c_p, c_p, C.POINTER(Grads), C.POINTER(PoolGrads), C.POINTER(PoolEmbeddingGrads), c_p]),
    'gsage_engine_backward_layer1_embedding': (C.c_int, [c_p, C.POINTER(EmbeddingGrads), c_p]),
    'gsage_engine_backward_layer1_linear': (C.c_int, [c_p, C.POINTER(LinearPrepGrads), c_p]),
    'gsage_wgrad': (C.c_int, [c_p, C.c_int, c_i64, C.c_int, c_p, C.c_int, c_i64, c_i64, c_p, C.c_int, c_i64, c_p, c_i64, C.c_int, c_p]),
    'gsage_adam_step': (C.c_int, [c_p, c_p, c_p, c_p, c_i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, c_i64, C.c_float, c_p, c_p]),
    'gsage_act_backward': (C.c_int, [c_p, c_i64, c_p, c_i64, c_i64, C.c_int, C.c_int, c_p, c_i64, c_p]),
    'gsage_segment_broadcast': (C.c_int, [c_p, c_i64, c_i64, C.c_int, C.c_int, C.c_float, c_p, c_i64, c_p]),
    'gsage_segment_max_backward': (C.c_int, [c_p, c_i64, c_p, c_i64, c_i64, C.c_int, C.c_int, c_p, c_i64, c_p]),
    'gsage_attention_sum_backward': (C.c_int, [c_p, c_i64, C.c_int, c_i64, C.c_int, c_p, c_i64, c_p, c_p, c_p, C.c_int, c_p, c_i64, c_p, c_p, c_p, c_p]),
    'gsage_colsum': (C.c_int, [c_p, c_i64, C.c_int, c_p, c_p]),
    'gsage_embedding_backward': (C.c_int, [c_p, c_i64, C.c_int, c_p, c_i64, c_p, c_i64, c_i64, c_p]),
    'gsage_l2_normalize_backward': (C.c_int, [c_p, c_p, c_i64, C.c_int, c_p, c_p]),
    'gsage_peer_allreduce_words': (c_i64, [c_i64]),
    'gsage_peer_allreduce': (C.c_int, [c_p, C.c_int, C.c_int, c_i64, C.c_uint32, C.c_float, c_p, c_p, c_p]),
    'gsage_metric_f1': (C.c_int, [c_p, c_i64, c_p, c_i64, c_i64, C.c_int, C.c_int, c_p, c_p, c_p]),
    'gsage_metric_mae': (C.c_int, [c_p, c_p, c_i64, c_p, c_p]),
    'gsage_engine_keep_activations': (C.c_int, [c_p, C.c_int]),
    'gsage_engine_profile': (C.c_int, [c_p, C.c_int]),
    'gsage_engine_profile_read': (C.c_int, [c_p, C.POINTER(C.c_double), C.POINTER(c_i64), C.POINTER(C.c_double), C.POINTER(C.c_double), c_p]),
}

EXPORTS = sorted(_SIGNATURES)

_lib = None


class GsageError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GsageError('libgsage_b200.so is not built (%s). Run `python -m pytorch_graphsage_b200.build` '
                             '-- there is no CPU fallback.' % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        if handle.gsage_abi_version() != ABI_VERSION:
            raise GsageError('libgsage_b200.so ABI mismatch')
        _lib = handle
    return _lib


def check(status):
    """Turn a negative gsage_status into the exception the reference would have raised."""
    if status == 0:
        return
    msg = lib().gsage_last_error().decode('utf-8', 'replace')
    if status == ERR_INDEX:
        raise IndexError(msg)
    if status == ERR_INVALID and 'n_samples must be set' in msg:
        raise AssertionError(msg)
    if status == ERR_INVALID:
        raise ValueError(msg)
    raise GsageError('gsage status %d: %s' % (status, msg))


def launch_count():
    return int(lib().gsage_launch_count())
