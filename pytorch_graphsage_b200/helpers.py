"""helpers.py of the reference (/root/reference/helpers.py:14-25): `set_seeds` seeds numpy's global stream, torch -- and the
device-resident MT19937 stream that continues numpy's; `to_numpy` is the tensor -> ndarray copy train.py uses for metrics
(the reference's own version recurses forever on torch >= 0.4)."""

from .rng import set_seeds          # noqa: F401


def to_numpy(x):
    return x.detach().cpu().numpy() if hasattr(x, 'detach') else x
