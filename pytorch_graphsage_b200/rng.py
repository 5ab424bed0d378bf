"""
DeviceMT19937 -- numpy's legacy global RandomState, resident on the GPU.

Host-side mirror of `helpers.set_seeds` (/root/reference/helpers.py:14-18) and of the two draws the
reference makes on the global stream (nn_modules.py:88 `np.random.choice`, problem.py:146
`np.random.permutation`).  `sync_to_numpy` / `sync_from_numpy` hand the exact stream position back and
forth, so host code that still calls np.random continues where the device stopped.
"""

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib
from . import ops


class DeviceMT19937(object):
    def __init__(self, seed=None):
        ops._bind_device()
        self._h = C.c_void_p()
        check(lib().gsage_rng_create(C.byref(self._h)))
        if seed is not None:
            self.seed(seed)

    def seed(self, seed):
        check(lib().gsage_rng_seed(self._h, int(seed) & 0xFFFFFFFF, ops.stream()))

    def get_state(self):
        """numpy-compatible ('MT19937', key, pos, 0, 0.0); synchronises the stream."""
        key = np.empty(624, dtype=np.uint32)
        pos = C.c_int()
        check(lib().gsage_rng_get_state(self._h, key.ctypes.data, C.byref(pos), ops.stream()))
        return ('MT19937', key, int(pos.value), 0, 0.0)

    def set_state(self, state):
        key = np.ascontiguousarray(state[1], dtype=np.uint32)
        check(lib().gsage_rng_set_state(self._h, key.ctypes.data, int(state[2]), ops.stream()))

    def sync_to_numpy(self):
        np.random.set_state(self.get_state())

    def sync_from_numpy(self):
        self.set_state(np.random.get_state())

    def raw(self, count):
        """Next `count` tempered 32-bit words as an int32-typed CUDA tensor (uint32 bit patterns)."""
        out = torch.empty((count,), dtype=torch.int32, device='cuda')
        check(lib().gsage_rng_raw(self._h, count, ops.ptr(out), ops.stream()))
        return out

    def randint(self, hi, count):
        """np.random.choice(hi, count) as an int32-typed CUDA tensor (uint32 bit patterns)."""
        out = torch.empty((count,), dtype=torch.int32, device='cuda')
        check(lib().gsage_rng_randint(self._h, int(hi), count, ops.ptr(out), ops.stream()))
        return out

    def permutation(self, n):
        """np.random.permutation(np.arange(n)) as an int64 CUDA tensor."""
        out = torch.empty((n,), dtype=torch.int64, device='cuda')
        check(lib().gsage_rng_permutation(self._h, n, ops.ptr(out), ops.stream()))
        return out

    def check(self):
        check(lib().gsage_rng_check(self._h, ops.stream()))

    def consumed(self):
        n = C.c_int64()
        check(lib().gsage_rng_consumed(self._h, C.byref(n), ops.stream()))
        return n.value

    def __del__(self):
        try:
            if self._h:
                lib().gsage_rng_destroy(self._h)
                self._h = None
        except Exception:
            pass


_default = [None]


def default_rng():
    """The process-wide device stream (the counterpart of numpy's global RandomState)."""
    if _default[0] is None:
        _default[0] = DeviceMT19937()
        _default[0].sync_from_numpy()
    return _default[0]


def set_seeds(seed=0):
    """helpers.set_seeds: numpy global, torch CPU/CUDA generators -- and the device stream."""
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        default_rng().seed(seed)
