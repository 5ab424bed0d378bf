"""Pins oracle/mt19937.py against numpy's legacy RandomState -- the object the reference calls
(/root/reference/nn_modules.py:88, helpers.py:15, problem.py:146)."""
import numpy as np
import pytest

from oracle.mt19937 import MT19937Oracle, init_genrand, mask_for


@pytest.mark.parametrize('seed', [0, 1, 5489, 123, 123 ** 2, 2 ** 32 - 1])
def test_seed_and_raw_words(seed):
    o, rs = MT19937Oracle(seed), np.random.RandomState(seed)
    st = rs.get_state()
    assert np.array_equal(o.key, st[1]) and o.pos == st[2] == 624
    assert np.array_equal(o.raw(2500), np.frombuffer(rs.bytes(10000), dtype='<u4'))
    st = rs.get_state()
    assert np.array_equal(o.key, st[1]) and o.pos == st[2]


def test_known_answer_5489():
    # first outputs of the MT19937 reference implementation for init_genrand(5489)
    assert list(MT19937Oracle(5489).raw(3)) == [3499211612, 581869302, 3890346734]
    assert init_genrand(5489)[0] == 5489


@pytest.mark.parametrize('hi', [1, 2, 3, 5, 37, 128, 129, 1000, 8763, 20000, 65536, 65537, 2 ** 31 - 1])
def test_bounded_matches_choice(hi):
    o, rs = MT19937Oracle(15129), np.random.RandomState(15129)
    for count in (1, 7, 640, 5000):
        assert np.array_equal(o.randint(hi, count), rs.choice(hi, count))
        st = rs.get_state()
        assert np.array_equal(o.key, st[1]) and o.pos == st[2], 'stream position diverged'


def test_hi_one_consumes_nothing():
    o = MT19937Oracle(9)
    before = (o.key.copy(), o.pos)
    assert not o.randint(1, 1000).any()
    assert np.array_equal(before[0], o.key) and before[1] == o.pos


@pytest.mark.parametrize('n', [1, 2, 10, 140, 3001])
def test_permutation(n):
    o, rs = MT19937Oracle(123), np.random.RandomState(123)
    assert np.array_equal(o.permutation(n), rs.permutation(np.arange(n)))
    assert np.array_equal(o.raw(5), np.frombuffer(rs.bytes(20), dtype='<u4'))


def test_state_handoff_roundtrip():
    rs = np.random.RandomState(77)
    rs.choice(1000, 321)
    o = MT19937Oracle()
    o.set_state(rs.get_state())
    a = o.randint(37, 100)
    rs2 = np.random.RandomState()
    rs2.set_state(o.get_state())
    assert np.array_equal(a, rs.choice(37, 100))
    assert np.array_equal(rs2.choice(5, 50), rs.choice(5, 50))


def test_mask():
    assert [mask_for(x) for x in (0, 1, 2, 3, 4, 127, 128, 19999)] == [0, 1, 3, 3, 7, 127, 255, 32767]
