"""
oracle/mt19937.py -- TEST INFRASTRUCTURE ONLY (not a product path).

CPU restatement of the random stream the reference's sparse sampler consumes.

The reference never implements an RNG itself: `SparseUniformNeighborSampler.__call__`
(/root/reference/nn_modules.py:88) calls `np.random.choice(maxdeg, (n, S))` on the *global*
numpy legacy `RandomState`, which `helpers.set_seeds` (/root/reference/helpers.py:14-18) seeds
with `np.random.seed(seed)`, and `NodeProblem.iterate` (/root/reference/problem.py:146) shares
the same stream through `np.random.permutation`.  The arithmetic therefore lives in a
third-party dependency, numpy's legacy MT19937 `RandomState` (version unpinned by the
reference; numpy's stream-compatibility policy freezes it; numpy 2.3.5 is what is installed
here and on the GPU box).  This file restates the published algorithm:

  * `init_genrand(s)`        -- Knuth-style seeding, 1812433253 multiplier (Matsumoto & Nishimura 2002)
  * the 624-word twist       -- x[n+624] = x[n+397] ^ twist(x[n], x[n+1])
  * tempering                -- the four shift/mask steps
  * legacy bounded integers  -- `randint(0, hi)` == masked rejection, one 32-bit word per attempt
  * legacy `permutation`     -- Fisher-Yates, i = n-1 .. 1, j = bounded(i)

It is pinned (tests/test_oracle_rng.py) against numpy's own `RandomState` -- raw words via
`RandomState.bytes`, bounded draws via `RandomState.choice`/`randint`, shuffles via
`RandomState.permutation`, and the state hand-off via `get_state`/`set_state` -- which is the
very object the reference calls.
"""

import numpy as np

N = 624
M = 397
MATRIX_A = np.uint32(0x9908B0DF)
UPPER = np.uint32(0x80000000)
LOWER = np.uint32(0x7FFFFFFF)


def init_genrand(seed):
    """mt[0] = s; mt[i] = 1812433253 * (mt[i-1] ^ (mt[i-1] >> 30)) + i   (mod 2^32)."""
    mt = np.zeros(N, dtype=np.uint64)
    mt[0] = np.uint64(seed & 0xFFFFFFFF)
    for i in range(1, N):
        prev = int(mt[i - 1])
        mt[i] = (1812433253 * (prev ^ (prev >> 30)) + i) & 0xFFFFFFFF
    return mt.astype(np.uint32)


def _mix(a, b):
    """twist(a, b): y = (a & UPPER) | (b & LOWER); (y >> 1) ^ (MATRIX_A if y odd)."""
    y = (a & UPPER) | (b & LOWER)
    return (y >> np.uint32(1)) ^ np.where((y & np.uint32(1)).astype(bool), MATRIX_A, np.uint32(0))


def twist(mt):
    """One full 624-word regeneration, as three dependency-free vector phases (227, 227, 170 words;
    the last word needs the *new* word 0)."""
    new = mt.copy()
    # phase 1: kk in [0, 227)  -> needs old[kk], old[kk+1], old[kk+397]
    new[0:227] = mt[397:624] ^ _mix(mt[0:227], mt[1:228])
    # phase 2: kk in [227, 454) -> needs new[kk-227], old[kk], old[kk+1]
    new[227:454] = new[0:227] ^ _mix(mt[227:454], mt[228:455])
    # phase 3: kk in [454, 623) -> needs new[kk-227]
    new[454:623] = new[227:396] ^ _mix(mt[454:623], mt[455:624])
    # last word wraps to the NEW word 0
    new[623:624] = new[396:397] ^ _mix(mt[623:624], new[0:1])
    return new


def temper(y):
    y = y.astype(np.uint32).copy()
    y ^= y >> np.uint32(11)
    y ^= (y << np.uint32(7)) & np.uint32(0x9D2C5680)
    y ^= (y << np.uint32(15)) & np.uint32(0xEFC60000)
    y ^= y >> np.uint32(18)
    return y


def mask_for(rng):
    """Smallest 2^k - 1 >= rng (numpy legacy `rk_interval` bit-smear)."""
    mask = int(rng)
    mask |= mask >> 1
    mask |= mask >> 2
    mask |= mask >> 4
    mask |= mask >> 8
    mask |= mask >> 16
    return mask


class MT19937Oracle(object):
    """Sequential legacy stream: `key` (624 untempered words) + `pos` (next word to hand out)."""

    def __init__(self, seed=None):
        if seed is not None:
            self.seed(seed)

    def seed(self, seed):
        self.key = init_genrand(int(seed))
        self.pos = N  # first draw triggers a twist (numpy: pos = RK_STATE_LEN after seeding)

    # -- state hand-off with numpy ------------------------------------------------------------
    def get_state(self):
        return ('MT19937', self.key.copy(), int(self.pos), 0, 0.0)

    def set_state(self, state):
        self.key = np.asarray(state[1], dtype=np.uint32).copy()
        self.pos = int(state[2])

    # -- raw words -----------------------------------------------------------------------------
    def raw(self, n):
        """Next n tempered 32-bit words."""
        out = np.empty(n, dtype=np.uint32)
        done = 0
        while done < n:
            if self.pos == N:
                self.key = twist(self.key)
                self.pos = 0
            take = min(n - done, N - self.pos)
            out[done:done + take] = temper(self.key[self.pos:self.pos + take])
            self.pos += take
            done += take
        return out

    # -- legacy bounded integers ---------------------------------------------------------------
    def randint(self, hi, count):
        """`RandomState.randint(0, hi, count)` == `choice(hi, count)` for hi <= 2^32:
        rng = hi-1; rng == 0 consumes no words; otherwise masked rejection, one word per attempt.
        Returns int64 values in output order."""
        rng = int(hi) - 1
        assert 0 <= rng <= 0xFFFFFFFF
        out = np.zeros(count, dtype=np.int64)
        if rng == 0 or count == 0:
            return out
        mask = np.uint32(mask_for(rng))
        done = 0
        while done < count:
            # draw a chunk, keep accepted values, give back the unused tail exactly
            want = count - done
            chunk = max(16, int(want * 1.25) + 16)
            save_key, save_pos = self.key.copy(), self.pos
            w = self.raw(chunk) & mask
            ok = np.flatnonzero(w <= np.uint32(rng))
            if ok.size >= want:
                used = int(ok[want - 1]) + 1           # raw words consumed
                out[done:] = w[ok[:want]]
                self.key, self.pos = save_key, save_pos
                self.raw(used)                         # re-advance exactly `used` words
                done = count
            else:
                out[done:done + ok.size] = w[ok]
                done += ok.size
        return out

    def permutation(self, n):
        """`RandomState.permutation(np.arange(n))` (Fisher-Yates from the top)."""
        arr = np.arange(n)
        for i in range(n - 1, 0, -1):
            j = int(self.randint(i + 1, 1)[0])
            arr[i], arr[j] = arr[j], arr[i]
        return arr
