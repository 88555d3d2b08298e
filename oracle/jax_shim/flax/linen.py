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
    """flax.linen.Module is a dataclass over the class annotations.  The shim has no variable collections: `self.param`
    returns the array the generator script registered under that name in `Module.shim_params`."""

    shim_params = {}

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        _dc.dataclass(cls)

    def param(self, name, init_fn, *shape_args):
        return _wrap(_np.asarray(Module.shim_params[name], _np.float32))


class _Initializers:
    """flax.linen.initializers: only looked up (he_normal / normal nodes are `call`ed), never used for values here."""

    @staticmethod
    def he_normal(*a, **k):
        return lambda *a_, **k_: None

    normal = zeros_init = lecun_normal = xavier_uniform = he_normal


initializers = _Initializers()


def relu(x):
    return _wrap(_np.maximum(_np.asarray(x), 0))


@_dc.dataclass
class Dropout:
    """flax.linen.Dropout: identity when deterministic (the only way the files we execute call it)."""
    rate: float = 0.0

    def __call__(self, x, deterministic=True):
        assert deterministic, "the shim has no RNG: stochastic dropout is not executed"
        return x


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
