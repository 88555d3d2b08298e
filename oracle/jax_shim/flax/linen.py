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
                    if hasattr(self, "setup") and not getattr(self, "_shim_setup_done", False):
                        # setup()-style module: children are attributes and take the attribute's name unless given one
                        object.__setattr__(self, "_shim_setup_done", True)
                        self.setup()
                        for attr, val in list(vars(self).items()):
                            if hasattr(val, "name") and getattr(val, "name", None) is None and not attr.startswith("_"):
                                try:
                                    object.__setattr__(val, "name", attr)
                                except Exception:
                                    pass
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

    normal = zeros_init = lecun_normal = xavier_uniform = variance_scaling = he_normal


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


# ---- image front end (tokenizers/images/image_tokenizer.py): Conv, GroupNorm, max_pool, gelu, Embed ------------------------
def gelu(x, approximate=True):
    """flax.linen.gelu = jax.nn.gelu, default approximate=True: 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))."""
    x = _np.asarray(x, _np.float32)
    if approximate:
        c = _np.float32(_np.sqrt(2.0 / _np.pi))
        return _wrap((0.5 * x * (1.0 + _np.tanh(c * (x + _np.float32(0.044715) * x * x * x)))).astype(_np.float32))
    from math import erf
    return _wrap((0.5 * x * (1.0 + _np.vectorize(erf)(x / _np.sqrt(2.0)))).astype(_np.float32))


def max_pool(inputs, window_shape, strides=None, padding="VALID"):
    """flax.linen.max_pool on [batch dims..., spatial..., features]: every leading dimension beyond the window's is batch."""
    assert padding == "VALID", "the shim pools without padding only"
    x = _np.asarray(inputs, _np.float32)
    wh, ww = window_shape
    sh, sw = strides or (1, 1)
    H, W = x.shape[-3], x.shape[-2]
    oh, ow = (H - wh) // sh + 1, (W - ww) // sw + 1
    out = _np.full(x.shape[:-3] + (oh, ow, x.shape[-1]), -_np.inf, _np.float32)
    for dy in range(wh):
        for dx in range(ww):
            out = _np.maximum(out, x[..., dy:dy + sh * oh:sh, dx:dx + sw * ow:sw, :])
    return _wrap(out)


@_dc.dataclass
class Conv:
    """flax.linen.Conv, 2-D, NHWC, kernel [kh, kw, in, features]; leading dimensions beyond (H, W, C) are batch (flattened and
    restored, as Flax does); padding VALID or SAME (SAME: total padding k - 1 for stride 1, low side (k - 1) // 2)."""
    features: int = 0
    kernel_size: object = (3, 3)
    strides: object = (1, 1)
    padding: str = "SAME"
    use_bias: bool = True
    dtype: object = None
    param_dtype: object = None
    kernel_init: object = None
    bias_init: object = None
    name: object = None

    def __call__(self, x):
        p = _child_params(self, "Conv")
        x = _np.asarray(x, _np.float32)
        lead = x.shape[:-3]
        x = x.reshape((-1,) + x.shape[-3:])
        k = _np.asarray(p["kernel"], _np.float32)
        kh, kw = k.shape[:2]
        sh, sw = self.strides
        if self.padding == "SAME":
            assert (sh, sw) == (1, 1)
            x = _np.pad(x, ((0, 0), ((kh - 1) // 2, kh - 1 - (kh - 1) // 2), ((kw - 1) // 2, kw - 1 - (kw - 1) // 2), (0, 0)))
        else:
            assert self.padding == "VALID"
        H, W = x.shape[1:3]
        oh, ow = (H - kh) // sh + 1, (W - kw) // sw + 1
        out = _np.zeros((x.shape[0], oh, ow, k.shape[3]), _np.float32)
        for dy in range(kh):
            for dx in range(kw):
                out += _np.einsum("bhwc,cf->bhwf", x[:, dy:dy + sh * oh:sh, dx:dx + sw * ow:sw, :], k[dy, dx], dtype=_np.float32)
        if self.use_bias:
            out = out + _np.asarray(p["bias"], _np.float32)
        assert out.shape[-1] == self.features
        return _wrap(out.reshape(lead + out.shape[1:]).astype(_np.float32))


@_dc.dataclass
class GroupNorm:
    """flax.linen.GroupNorm (0.8.x): reduction_axes default = every axis except the FIRST (batch) one, i.e. on a
    [B, images, patches, H, W, C] input the statistics of a group run over images x patches x H x W x (C / groups) for each
    batch row; use_fast_variance (var = max(0, E[x^2] - E[x]^2)); scale / bias per channel."""
    num_groups: object = 32
    group_size: object = None
    epsilon: float = 1e-6
    dtype: object = None
    param_dtype: object = None
    use_bias: bool = True
    use_scale: bool = True
    name: object = None

    def __call__(self, x):
        p = _child_params(self, "GroupNorm")
        x = _np.asarray(x, _np.float32)
        C = x.shape[-1]
        G = self.num_groups if self.num_groups is not None else C // self.group_size
        assert C % G == 0
        g = x.reshape(x.shape[:-1] + (G, C // G))
        ax = tuple(range(1, g.ndim - 2)) + (g.ndim - 1,)
        mu = g.mean(axis=ax, keepdims=True, dtype=_np.float32)
        var = _np.maximum((g * g).mean(axis=ax, keepdims=True, dtype=_np.float32) - mu * mu, 0)
        y = ((g - mu) * (1.0 / _np.sqrt(var + _np.float32(self.epsilon))).astype(_np.float32)).reshape(x.shape)
        if self.use_scale:
            y = y * _np.asarray(p["scale"], _np.float32)
        if self.use_bias:
            y = y + _np.asarray(p["bias"], _np.float32)
        return _wrap(y.astype(_np.float32))


@_dc.dataclass
class Embed:
    """flax.linen.Embed: rows of the [num_embeddings, features] table."""
    num_embeddings: int = 0
    features: int = 0
    dtype: object = None
    param_dtype: object = None
    embedding_init: object = None
    name: object = None

    def __call__(self, idx):
        p = _child_params(self, "Embed")
        return _wrap(_np.asarray(p["embedding"], _np.float32)[_np.asarray(idx, _np.int64)])
