"""CPU-only trace of xla_ffi/tome_jax.py against a stub of the `jax` symbols it uses.

jax / jaxlib are not installed in this image, so the jax.ffi layer cannot be compiled or run here.  This test only
proves what can be proved without them: the module imports, `mode` / `r` / flags reach `ffi_call` as static numpy
attributes (never as traced positionals -- ADVICE r1), the custom_vjp forward returns the (y, size) pair `merge()` and
`merge_wavg()` index, and every declared result shape follows token_compression.py:54-129."""
import ctypes
import importlib
import sys
import types

import numpy as np
import pytest


class _Tracer:
    """stands in for a traced array: anything that tries to turn it into a Python/numpy scalar raises, as under jit"""

    def __init__(self, a):
        self.a = np.asarray(a)
        self.shape, self.dtype = self.a.shape, self.a.dtype

    def __getitem__(self, i):
        return _Tracer(self.a[i])

    def astype(self, dt):
        return _Tracer(self.a.astype(dt))

    def __int__(self):
        raise TypeError("tracer converted to a concrete int")

    __index__ = __float__ = __int__


def _stub_jax(calls):
    jax = types.ModuleType("jax")
    jnp = types.ModuleType("jax.numpy")
    jnp.float32, jnp.int32 = np.float32, np.int32
    jnp.ones = lambda shape, dtype=np.float32: _Tracer(np.ones(shape, dtype))
    jnp.zeros_like = lambda t: _Tracer(np.zeros(t.shape, t.dtype))
    ffi = types.SimpleNamespace()
    ffi.register_ffi_target = lambda name, capsule, platform=None: calls.setdefault("registered", []).append(name)
    ffi.pycapsule = lambda sym: sym

    def ffi_call(name, result):
        def run(*operands, **attrs):
            for k, v in attrs.items():
                assert isinstance(v, np.generic), f"{name}: attribute {k} must be a static numpy scalar, got {type(v)}"
            for o in operands:
                assert isinstance(o, _Tracer), f"{name}: operand of type {type(o)}"
            calls.setdefault("ffi", []).append((name, attrs))
            mk = lambda s: _Tracer(np.zeros(s.shape, s.dtype))  # noqa: E731
            return tuple(mk(s) for s in result) if isinstance(result, tuple) else mk(result)
        return run

    ffi.ffi_call = ffi_call
    jax.ffi = ffi
    jax.ShapeDtypeStruct = lambda shape, dtype: types.SimpleNamespace(shape=tuple(shape), dtype=np.dtype(dtype))

    class custom_vjp:  # noqa: N801
        def __init__(self, f):
            self.f = f

        def defvjp(self, fwd, bwd):
            self.fwd, self.bwd = fwd, bwd

        def __call__(self, *args):
            out, res = self.fwd(*args)          # exercise the residual plumbing too
            grads = self.bwd(res, out)
            assert len(grads) == len(args)
            return out

    jax.custom_vjp = custom_vjp
    jax.numpy = jnp
    return jax, jnp


@pytest.fixture()
def tome_jax(monkeypatch):
    calls = {}
    jax, jnp = _stub_jax(calls)
    monkeypatch.setitem(sys.modules, "jax", jax)
    monkeypatch.setitem(sys.modules, "jax.numpy", jnp)
    real_cdll = ctypes.CDLL

    def fake_cdll(path, *a, **k):
        if str(path).endswith("libtome_xla_ffi.so"):
            class _L:
                def __getattr__(self, n):
                    return n
            return _L()
        return real_cdll(path, *a, **k)

    monkeypatch.setattr(ctypes, "CDLL", fake_cdll)
    name = "multi_modal_transformers_tokenmerge_b200.xla_ffi.tome_jax"
    sys.modules.pop(name, None)
    mod = importlib.import_module(name)
    yield mod, calls
    sys.modules.pop(name, None)


def test_matching_and_merge_trace(tome_jax):
    mod, calls = tome_jax
    assert {"tome_sim_argmax", "tome_select_topr", "tome_merge_fwd", "tome_merge_bwd"} <= set(calls["registered"])
    B, T, C, r = 2, 11, 8, 3
    metric = _Tracer(np.zeros((B, T, 4), np.float32))
    merge = mod.bipartite_soft_matching(metric, r)
    assert merge.r == r
    x = _Tracer(np.zeros((B, T, C), np.float32))
    y = merge(x)                                   # mode="sum"
    assert y.shape == (B, T - r, C)
    with pytest.raises(ValueError):
        merge(x, mode="mean")                      # the reference implements "sum" only (token_compression.py:99)
    y, size = mod.merge_wavg(merge, x)
    assert y.shape == (B, T - r, C) and size.shape == (B, T - r, 1)
    modes = [int(a["mode"]) for n, a in calls["ffi"] if n == "tome_merge_fwd"]
    assert modes == [0, 1]
    bwd_modes = [int(a["mode"]) for n, a in calls["ffi"] if n == "tome_merge_bwd"]
    assert bwd_modes == [0, 1]


def test_r_clamps_to_identity(tome_jax):
    mod, _ = tome_jax
    metric = _Tracer(np.zeros((1, 4, 4), np.float32))
    merge = mod.bipartite_soft_matching(metric, 0)
    x = _Tracer(np.zeros((1, 4, 8), np.float32))
    assert merge(x) is x and merge.r == 0
    y, size = mod.merge_wavg(merge, x)
    assert y is x and size.shape == (1, 4, 1)
