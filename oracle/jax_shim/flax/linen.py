import dataclasses as _dc

import numpy as _np
from jax.numpy import _wrap


def make_causal_mask(x, extra_batch_dims=0, dtype=_np.float32):
    """flax.linen.make_causal_mask: [..., 1, L, L] lower-triangular (q >= k) mask of ones."""
    n = _np.shape(x)[-1]
    idx = _np.arange(n)
    m = (idx[:, None] >= idx[None, :]).astype(dtype)
    return _wrap(m.reshape((1,) * (_np.ndim(x) - 1) + (1, n, n)))


# ---- a minimal stand-in for Flax's variable scopes -------------------------------------------------------------------------
# A scope = (parameter subtree, per-class counters).  A module called inside a compact method takes the next auto-name of
# its class (`LayerNorm_0`, `LayerNorm_1`, `Dense_0`, ... -- Flax's naming rule) or its explicit `name=`, and finds its
# parameters under that key of the enclosing scope's subtree.  Only used when the generator script installs a parameter
# tree with `shim_scope(tree)`; the older fixtures pass parameters through the config nodes instead and never open a scope.
_scopes = []


class shim_scope:
    def __init__(self, params):
        self.params, self.counters = params, {}

    def __enter__(self):
        _scopes.append(self)
        return self

    def __exit__(self, *a):
        _scopes.pop()


def _child_params(obj, cls_name):
    """parameter subtree of a module instance that is being called in the current scope"""
    sc = _scopes[-1]
    forced = getattr(obj, "_shim_params", None)
    if forced is not None:                      # nn.scan hands the sliced subtree to the scanned module directly
        return forced
    name = getattr(obj, "name", None)
    if name is None:
        i = sc.counters.get(cls_name, 0)
        sc.counters[cls_name] = i + 1
        name = f"{cls_name}_{i}"
    return sc.params[name]


def merge_param(name, a, b):
    """flax.linen.merge_param: exactly one of the module attribute and the call argument is given"""
    if a is None and b is None:
        raise ValueError(f"Parameter \"{name}\" must be passed to the constructor or at call time.")
    if a is not None and b is not None:
        raise ValueError(f"Parameter \"{name}\" was passed to the constructor and at call time.")
    return b if a is None else a


class Module:
    """flax.linen.Module is a dataclass over the class annotations.  Without an open shim_scope `self.param` returns the
    array the generator script registered under that name in `Module.shim_params`; inside one, the array under that name
    of the module's own parameter subtree, and calling a module opens the scope of its children."""

    shim_params = {}

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        if "name" not in getattr(cls, "__annotations__", {}):
            cls.__annotations__ = dict(getattr(cls, "__annotations__", {}), name=object)
            cls.name = None
        _dc.dataclass(cls)
        call = cls.__dict__.get("__call__")
        if call is not None:
            def wrapped(self, *a, __call=call, **k):
                if not _scopes:
                    return __call(self, *a, **k)
                with shim_scope(_child_params(self, type(self).__name__)) as sc:
                    self._shim_scope = sc
                    return __call(self, *a, **k)
            cls.__call__ = wrapped

    def param(self, name, init_fn, *shape_args):
        if _scopes:
            return _wrap(_np.asarray(self._shim_scope.params[name], _np.float32))
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

    name: object = None

    def __call__(self, x):
        kernel, bias = self.kernel, self.bias
        if kernel is None:   # parameters from the enclosing scope (auto-name Dense_i), as Flax would
            p = _child_params(self, "Dense")
            kernel, bias = p["kernel"], p.get("bias")
        y = _np.matmul(_np.asarray(x, _np.float32), _np.asarray(kernel, _np.float32))
        assert y.shape[-1] == self.features
        if self.use_bias:
            y = y + _np.asarray(bias, _np.float32)
        return _wrap(y)


@_dc.dataclass
class LayerNorm:
    """flax.linen.LayerNorm (0.8.x): statistics over `reduction_axes`, scale / bias over `feature_axes`,
    use_fast_variance: var = max(0, E[x^2] - E[x]^2); y = (x - mean) * rsqrt(var + epsilon) * scale + bias."""
    epsilon: float = 1e-6
    dtype: object = None
    param_dtype: object = None
    use_bias: bool = True
    use_scale: bool = True
    reduction_axes: object = -1
    feature_axes: object = -1
    use_fast_variance: bool = True
    name: object = None

    def __call__(self, x):
        p = _child_params(self, "LayerNorm")
        x = _np.asarray(x, _np.float32)
        ax = tuple(a % x.ndim for a in ([self.reduction_axes] if isinstance(self.reduction_axes, int) else self.reduction_axes))
        fa = tuple(a % x.ndim for a in ([self.feature_axes] if isinstance(self.feature_axes, int) else self.feature_axes))
        mu = x.mean(axis=ax, keepdims=True, dtype=_np.float32)
        var = _np.maximum((x * x).mean(axis=ax, keepdims=True, dtype=_np.float32) - mu * mu, 0) if self.use_fast_variance \
            else ((x - mu) ** 2).mean(axis=ax, keepdims=True, dtype=_np.float32)
        y = (x - mu) * (1.0 / _np.sqrt(var + _np.float32(self.epsilon))).astype(_np.float32)
        shape = [x.shape[i] if i in fa else 1 for i in range(x.ndim)]
        if self.use_scale:
            y = y * _np.asarray(p["scale"], _np.float32).reshape(shape)
        if self.use_bias:
            y = y + _np.asarray(p["bias"], _np.float32).reshape(shape)
        return _wrap(y)


@_dc.dataclass
class MultiHeadDotProductAttention:
    """flax.linen.MultiHeadDotProductAttention / SelfAttention (0.8.x), deterministic path: DenseGeneral query / key / value
    ([C, H, D] kernels, [H, D] biases), dot_product_attention (query scaled by 1/sqrt(D), mask -> finfo.min, softmax),
    DenseGeneral out ([H, D, C] kernel).  Accepts the reference's call `attn(x, x, mask=mask, deterministic=...)`."""
    num_heads: int = 1
    qkv_features: object = None
    out_features: object = None
    dropout_rate: float = 0.0
    broadcast_dropout: bool = True
    decode: bool = False
    use_bias: bool = True
    normalize_qk: bool = False
    dtype: object = None
    param_dtype: object = None
    kernel_init: object = None
    bias_init: object = None
    name: object = None

    def __call__(self, inputs_q, inputs_kv=None, mask=None, deterministic=None):
        assert deterministic or self.dropout_rate == 0.0, "the shim has no RNG: stochastic dropout is not executed"
        assert not self.decode and not self.normalize_qk
        p = _child_params(self, type(self).__name__)
        xq = _np.asarray(inputs_q, _np.float32)
        xkv = xq if inputs_kv is None else _np.asarray(inputs_kv, _np.float32)
        proj = lambda x, n: _np.einsum("btc,chd->bthd", x, _np.asarray(p[n]["kernel"], _np.float32)) + _np.asarray(p[n]["bias"], _np.float32)  # noqa: E731
        q, k, v = proj(xq, "query"), proj(xkv, "key"), proj(xkv, "value")
        d = q.shape[-1]
        logits = _np.einsum("bqhd,bkhd->bhqk", q / _np.sqrt(_np.float32(d)), k)
        if mask is not None:
            logits = _np.where(_np.asarray(mask, bool), logits, _np.finfo(_np.float32).min)
        w = _np.exp(logits - logits.max(-1, keepdims=True))
        w = w / w.sum(-1, keepdims=True)
        o = _np.einsum("bhqk,bkhd->bqhd", w.astype(_np.float32), v)
        out = _np.einsum("bthd,hdc->btc", o, _np.asarray(p["out"]["kernel"], _np.float32)) + _np.asarray(p["out"]["bias"], _np.float32)
        return _wrap(out.astype(_np.float32))


class SelfAttention(MultiHeadDotProductAttention):
    pass


def scan(target, variable_axes=None, variable_broadcast=False, split_rngs=None, length=None, **_):
    """flax.linen.scan over a Module class with variable_axes={'params': 0}: the scanned module's parameters carry a leading
    [length] axis and live under the auto-name `Scan<Class>_i`; iteration i runs the module on slice i, threading the carry."""
    assert variable_axes == {"params": 0}

    def factory(**kw):
        class _Scanned:
            name = None

            def __call__(self_, carry, xs):
                stacked = _child_params(self_, "Scan" + target.__name__)
                ys = None
                for i in range(length):
                    inner = target(**kw)
                    inner._shim_params = _tree_index(stacked, i)
                    carry, ys = inner(carry, xs)
                return carry, ys
        return _Scanned()
    return factory


def _tree_index(tree, i):
    return {k: _tree_index(v, i) for k, v in tree.items()} if isinstance(tree, dict) else _np.asarray(tree)[i]
