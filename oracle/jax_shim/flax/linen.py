import dataclasses as _dc

import numpy as _np
from jax.numpy import _wrap


def make_causal_mask(x, extra_batch_dims=0, dtype=_np.float32):
    """flax.linen.make_causal_mask: [..., 1, L, L] lower-triangular (q >= k) mask of ones."""
    n = _np.shape(x)[-1]
    idx = _np.arange(n)
    m = (idx[:, None] >= idx[None, :]).astype(dtype)
    return _wrap(m.reshape((1,) * (_np.ndim(x) - 1) + (1, n, n)))


class Module:
    """flax.linen.Module is a dataclass over the class annotations; that is all the action heads need."""

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        _dc.dataclass(cls)


def compact(fn):
    return fn


@_dc.dataclass
class Dense:
    """flax.linen.Dense: y = x @ kernel + bias.  The shim has no parameter store: the generator script passes the
    kernel [in, features] and bias [features] through the config node itself."""
    features: int
    kernel: object = None
    bias: object = None
    use_bias: bool = True
    kernel_init: object = None
    bias_init: object = None

    def __call__(self, x):
        y = _np.matmul(_np.asarray(x, _np.float32), _np.asarray(self.kernel, _np.float32))
        assert y.shape[-1] == self.features
        if self.use_bias:
            y = y + _np.asarray(self.bias, _np.float32)
        return _wrap(y)
