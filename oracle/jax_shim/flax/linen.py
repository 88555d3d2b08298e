import numpy as _np
from jax.numpy import _wrap


def make_causal_mask(x, extra_batch_dims=0, dtype=_np.float32):
    """flax.linen.make_causal_mask: [..., 1, L, L] lower-triangular (q >= k) mask of ones."""
    n = _np.shape(x)[-1]
    idx = _np.arange(n)
    m = (idx[:, None] >= idx[None, :]).astype(dtype)
    return _wrap(m.reshape((1,) * (_np.ndim(x) - 1) + (1, n, n)))


class Module:  # never instantiated by the files we execute
    pass
