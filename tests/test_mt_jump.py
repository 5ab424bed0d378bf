"""CPU pin of the MT19937 jump-ahead polynomials (csrc/mt_jump.cpp) against numpy's own stream: jumping a
state by J words with g_J(t) = t^J mod phi(t) must land exactly where numpy lands after drawing J words."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope='module')
def lib():
    from pytorch_graphsage_b200 import build
    h = C.CDLL(build.build())
    h.gsage_mt_jump_poly.argtypes = [C.c_uint64, C.c_void_p]
    h.gsage_mt_jump_poly.restype = C.c_int
    h.gsage_mt_jump_apply_host.argtypes = [C.c_void_p] * 3
    h.gsage_mt_jump_apply_host.restype = None
    return h


@pytest.mark.parametrize('blocks', [1, 2, 33, 256, 256 * 7, 256 * 31])
def test_jump_matches_sequential_stream(lib, blocks):
    J = 624 * blocks
    poly = np.zeros(624, dtype=np.uint32)
    assert lib.gsage_mt_jump_poly(J, poly.ctypes.data) == 0
    for seed in (1, 15129):
        rs = np.random.RandomState(seed)
        rs.bytes(4 * 624)                                   # first twist: key is now stream block 0
        key = rs.get_state()[1].copy()
        out = np.zeros(624, dtype=np.uint32)
        lib.gsage_mt_jump_apply_host(key.ctypes.data, poly.ctypes.data, out.ctypes.data)
        rs.bytes(4 * J)
        want = rs.get_state()[1]
        # the low 31 bits of the first word are not part of the 19937-bit state
        assert np.array_equal(out[1:], want[1:]) and (out[0] >> 31) == (want[0] >> 31)


def test_polynomial_degree_below_19937(lib):
    poly = np.zeros(624, dtype=np.uint32)
    assert lib.gsage_mt_jump_poly(624 * 256 * 5, poly.ctypes.data) == 0
    bits = np.unpackbits(poly.view(np.uint8), bitorder='little')
    assert not bits[19937:].any() and 8000 < bits.sum() < 12000
