"""jax.ffi registration + custom_vjp wrappers for the handlers of tome_xla_ffi.cc.

NEVER RUN AGAINST REAL JAX IN THIS IMAGE (jax / jaxlib are not installed): importing this module without
JAX raises ImportError at once.  tests/test_xla_ffi_stub.py traces it on the CPU against a stub of the few `jax` symbols it
uses (ffi_call returning zero arrays of the declared shapes), which checks syntax, static-attribute handling and shapes --
not the kernels and not XLA.  It shows exactly what a maintainer of the reference adds so that
`multi_modal_transformers.tokenizers.token_compression.merge_wavg` (and the block modules built on it) run on
libtome_b200.so from inside jit-compiled Flax code.  Logic-free: shapes in, `jax.ffi.ffi_call`, shapes out.
"""
from __future__ import annotations

import ctypes
import os

import jax            # noqa: F401  (ImportError here is the intended failure mode without JAX)
import jax.numpy as jnp
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = ctypes.CDLL(os.path.join(_HERE, "libtome_xla_ffi.so"))
for _name, _sym in [("tome_sim_argmax", "TomeSimArgmax"), ("tome_select_topr", "TomeSelectTopR"), ("tome_merge_fwd", "TomeMergeFwd"),
                    ("tome_merge_bwd", "TomeMergeBwd"), ("tome_attention_fwd", "TomeAttentionFwd"),
                    ("tome_attention_bwd", "TomeAttentionBwd"), ("tome_dense", "TomeDense"), ("tome_layernorm_fwd", "TomeLayerNormFwd"),
                    ("tome_stack_fwd", "TomeStackFwd")]:
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_lib, _sym)), platform="CUDA")

S = jax.ShapeDtypeStruct


def bipartite_soft_matching(metric, r, class_token=False, distill_token=False):
    """token_compression.py:54-112 -> `merge` closure over device-resident indices."""
    b, t, _ = metric.shape
    r = max(0, min(r, (t - int(class_token) - int(distill_token)) // 2))
    ta, tb = (t + 1) // 2, t // 2
    if r == 0:
        ident = lambda x, mode="sum": x  # noqa: E731  (the reference returns a tuple here, :70)
        ident.r = 0
        return ident
    nmax, nidx = jax.ffi.ffi_call("tome_sim_argmax", (S((b, ta), jnp.float32), S((b, ta), jnp.int32)))(
        metric, class_token=np.int32(class_token), distill_token=np.int32(distill_token))
    edge, dst, row_map, dst_off, dst_src = jax.ffi.ffi_call(
        "tome_select_topr", (S((b, ta), jnp.int32), S((b, r), jnp.int32), S((b, t), jnp.int32), S((b, tb + 1), jnp.int32),
                             S((b, r), jnp.int32)))(nmax, nidx, tokens=np.int32(t), r=np.int32(r), distill_token=np.int32(distill_token))

    def _make_merge(mode: int):
        """One custom_vjp per merge mode: `mode` is a Python int closed over (a static FFI attribute), never a traced
        argument -- as a positional of the custom_vjp it would reach `_fwd` as a tracer and `np.int32(mode)` would raise."""

        @jax.custom_vjp
        def f(x, size):
            return _fwd(x, size)[0]

        def _fwd(x, size):
            c = x.shape[-1]
            y, s_out = jax.ffi.ffi_call("tome_merge_fwd", (S((b, t - r, c), x.dtype), S((b, t - r), jnp.float32)))(
                x, size, edge, dst_off, dst_src, r=np.int32(r), mode=np.int32(mode), distill_token=np.int32(distill_token))
            return (y, s_out), (size, s_out)

        def _bwd(res, g):
            size, s_out = res
            dy, _ = g
            dx = jax.ffi.ffi_call("tome_merge_bwd", S((b, t, dy.shape[-1]), dy.dtype))(
                dy, size, s_out, row_map, r=np.int32(r), mode=np.int32(mode))
            return dx, jnp.zeros_like(size)   # sizes descend from constants only: no gradient

        f.defvjp(_fwd, _bwd)
        return f

    _merge_sum, _merge_wavg = _make_merge(0), _make_merge(1)

    def merge(x, mode="sum"):
        """merge(x, mode="sum") of :90-109 (fp32 rows of >= 4 columns; merge_wavg below never needs the 1-column call)."""
        if mode != "sum":
            raise ValueError(f'merge: mode "{mode}" is not implemented (the reference only implements "sum", :99)')
        y, _ = _merge_sum(x, jnp.ones(x.shape[:2], jnp.float32))
        return y

    merge.r, merge.wavg = r, _merge_wavg   # wavg(x, size [B,T]) -> (y, size_out [B,T-r])
    return merge


def merge_wavg(merge, x, size=None):
    """token_compression.py:114-129 in one kernel pass; returns (x [B,T-r,C], size [B,T-r,1])."""
    if getattr(merge, "r", 0) == 0:
        return x, (jnp.ones(x.shape[:2] + (1,), jnp.float32) if size is None else size)
    s = jnp.ones(x.shape[:2], jnp.float32) if size is None else size[..., 0].astype(jnp.float32)
    y, s_out = merge.wavg(x, s)
    return y, s_out[..., None]
