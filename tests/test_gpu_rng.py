"""Device MT19937 stream (C ABI gsage_rng_*) vs numpy's legacy RandomState -- bit-exact, including the stream
position hand-off.  GPU only."""
import numpy as np
import pytest
import torch

from oracle.mt19937 import MT19937Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def g():
    import pytorch_graphsage_b200 as g
    return g


@pytest.mark.parametrize('seed', [0, 5489, 123 ** 2, 2 ** 32 - 1])
def test_raw_words(g, seed):
    r = g.DeviceMT19937(seed)
    rs = np.random.RandomState(seed)
    for count in (1, 623, 624, 625, 100000):
        got = g.ops.u32_to_numpy(r.raw(count))
        want = np.frombuffer(rs.bytes(4 * count), dtype='<u4')
        assert np.array_equal(got, want)
    st, want = r.get_state(), rs.get_state()
    assert np.array_equal(st[1], want[1]) and st[2] == want[2]


def test_seed_state_matches_numpy(g):
    r = g.DeviceMT19937(777)
    st, want = r.get_state(), np.random.RandomState(777).get_state()
    assert np.array_equal(st[1], want[1]) and st[2] == want[2] == 624


@pytest.mark.parametrize('hi', [1, 2, 3, 37, 128, 129, 1000, 8763, 20000, 65537, 2 ** 31 - 1])
def test_bounded_draws(g, hi):
    r = g.DeviceMT19937(15129)
    rs = np.random.RandomState(15129)
    for count in (1, 7, 2047, 2048, 2049, 140800):
        got = g.ops.u32_to_numpy(r.randint(hi, count)).astype(np.int64)
        assert np.array_equal(got, rs.choice(hi, count)), (hi, count)
    st, want = r.get_state(), rs.get_state()
    assert np.array_equal(st[1], want[1]) and st[2] == want[2], 'stream position diverged'
    r.check()


def test_oracle_agrees_too(g):
    r, o = g.DeviceMT19937(99), MT19937Oracle(99)
    assert np.array_equal(g.ops.u32_to_numpy(r.randint(777, 5000)).astype(np.int64), o.randint(777, 5000))
    st = r.get_state()
    assert np.array_equal(st[1], o.key) and st[2] == o.pos


def test_long_stream_crosses_ring_and_resync(g):
    """More words than the ring holds: generation must wrap and the host bounds must re-tighten."""
    import os
    os.environ['GSAGE_RNG_LOG2_WORDS'] = '16'        # 65536-word ring
    try:
        r = g.DeviceMT19937(4242)
    finally:
        del os.environ['GSAGE_RNG_LOG2_WORDS']
    rs = np.random.RandomState(4242)
    for it in range(12):
        count = 30000 + 1111 * it
        got = g.ops.u32_to_numpy(r.randint(20000, count)).astype(np.int64)
        assert np.array_equal(got, rs.choice(20000, count)), it
    got = g.ops.u32_to_numpy(r.raw(200000))            # one call larger than the ring: split in pieces
    assert np.array_equal(got, np.frombuffer(rs.bytes(800000), dtype='<u4'))
    st, want = r.get_state(), rs.get_state()
    assert np.array_equal(st[1], want[1]) and st[2] == want[2]


def test_state_handoff_both_ways(g):
    rs = np.random.RandomState(31337)
    rs.choice(1000, 321)
    r = g.DeviceMT19937()
    r.set_state(rs.get_state())
    a = g.ops.u32_to_numpy(r.randint(37, 1000)).astype(np.int64)
    assert np.array_equal(a, rs.choice(37, 1000))
    rs2 = np.random.RandomState()
    rs2.set_state(r.get_state())
    assert np.array_equal(rs2.choice(5, 50), rs.choice(5, 50))


def test_global_stream_sync(g):
    g.set_seeds(123)
    dev = g.default_rng()
    a = g.ops.u32_to_numpy(dev.randint(1000, 10)).astype(np.int64)
    dev.sync_to_numpy()
    b = np.random.choice(1000, 10)
    ref = np.random.RandomState(123)
    assert np.array_equal(a, ref.choice(1000, 10)) and np.array_equal(b, ref.choice(1000, 10))
    np.random.choice(77, 5)
    dev.sync_from_numpy()
    ref.choice(77, 5)
    assert np.array_equal(g.ops.u32_to_numpy(dev.randint(9, 4)).astype(np.int64), ref.choice(9, 4))


@pytest.mark.parametrize('n', [1, 2, 10, 140, 5000])
def test_permutation(g, n):
    r, rs = g.DeviceMT19937(123), np.random.RandomState(123)
    assert np.array_equal(r.permutation(n).cpu().numpy(), rs.permutation(np.arange(n)))
    assert np.array_equal(g.ops.u32_to_numpy(r.raw(5)), np.frombuffer(rs.bytes(20), dtype='<u4'))


def test_lane_parallel_refill_matches_numpy(g):
    """Big requests go through the jump-ahead lane refill (32 lanes x 256 blocks): still the same stream."""
    r, rs = g.DeviceMT19937(2024), np.random.RandomState(2024)
    got = g.ops.u32_to_numpy(r.raw(7000000))                       # > one 5.1 M-word lane refill
    assert np.array_equal(got, np.frombuffer(rs.bytes(28000000), dtype='<u4'))
    for hi, count in ((20000, 2252800), (129, 1000000), (20000, 204800)):
        got = g.ops.u32_to_numpy(r.randint(hi, count)).astype(np.int64)
        assert np.array_equal(got, rs.choice(hi, count)), (hi, count)
    st, want = r.get_state(), rs.get_state()
    assert np.array_equal(st[1], want[1]) and st[2] == want[2]
    r.check()
