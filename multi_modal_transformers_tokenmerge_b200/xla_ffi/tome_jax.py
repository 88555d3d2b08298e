"""jax.ffi registration + custom_vjp wrappers for the handlers of tome_xla_ffi.cc.

NOT EXERCISED IN THIS IMAGE (jax / jaxlib are not installed): importing this module without JAX raises
ImportError at once.  It shows exactly what a maintainer of the reference adds so that
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

    @jax.custom_vjp
    def _merge(x, size, mode):
        return _fwd(x, size, mode)[0]

    def _fwd(x, size, mode):
        c = x.shape[-1]
        y, s_out = jax.ffi.ffi_call("tome_merge_fwd", (S((b, t - r, c), x.dtype), S((b, t - r), jnp.float32)))(
            x, size, edge, dst_off, dst_src, r=np.int32(r), mode=np.int32(mode), distill_token=np.int32(distill_token))
        return (y, s_out), (size, s_out, mode)

    def _bwd(res, g):
        size, s_out, mode = res
        dy, _ = g
        dx = jax.ffi.ffi_call("tome_merge_bwd", S((b, t, dy.shape[-1]), dy.dtype))(dy, size, s_out, row_map, r=np.int32(r), mode=np.int32(mode))
        return dx, None, None

    _merge.defvjp(_fwd, _bwd)
    _merge = jax.tree_util.Partial(_merge)

    def merge(x, mode="sum"):
        """merge(x, mode="sum") of :90-109 (fp32 rows of >= 4 columns; merge_wavg below never needs the 1-column call)."""
        if mode != "sum":
            raise ValueError(f'merge: mode "{mode}" is not implemented (the reference only implements "sum", :99)')
        return _merge(x, jnp.ones(x.shape[:2], jnp.float32), 0)[0]

    merge.r, merge.wavg = r, lambda x, size: _merge(x, size, 1)
    return merge


def merge_wavg(merge, x, size=None):
    """token_compression.py:114-129 in one kernel pass; returns (x [B,T-r,C], size [B,T-r,1])."""
    if getattr(merge, "r", 0) == 0:
        return x, (jnp.ones(x.shape[:2] + (1,), jnp.float32) if size is None else size)
    s = jnp.ones(x.shape[:2], jnp.float32) if size is None else size[..., 0].astype(jnp.float32)
    y, s_out = merge.wavg(x, s)
    return y, s_out[..., None]
