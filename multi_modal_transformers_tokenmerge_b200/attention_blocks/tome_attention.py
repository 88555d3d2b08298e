"""Attention block with ToMe: the module API of multi_modal_transformers/attention_blocks/tome_attention.py on the
sm_100a kernels.

    ToMeMultiHeadDotProductAttention(num_heads, dtype, param_dtype, qkv_features, out_features, broadcast_dropout,
        dropout_rate, deterministic, precision, kernel_init, bias_init, use_bias, attention_fn, decode, normalize_qk)
        (inputs_q, inputs_k=None, inputs_v=None, *, inputs_kv=None, mask=None, deterministic=None,
         dropout_rng=None, sow_weights=False)                                                    :19-300
    ToMeEncoder1DBlock(layer_norm, dropout, self_attention, mlp_block, train, mask)(inputs, mask, train)   :305-333
    AddPositionEmbedding(posemb_init)(inputs)                                                    :335-349
    StackedEncoder1DBlock(num_blocks, encoder_1d_block)(x, train=False, mask=None)               :351-383

The reference file does not parse and its ToMe step is a five-line stub (:249-256); what is built
here is its stated intent with the ToMe-paper placement:

    x = x + attn(LN(x), block-causal mask, bias = log size)
    metric = keys reduced over heads -> bipartite_soft_matching -> x, size = merge_wavg(merge, x, size)
    x = x + MLP(LN(x))

New knobs are additive attributes with reference-literal defaults: `tome_r` (tokens merged per block; the stub
hard-codes r = 5 at :252, default here 0 = plain attention), `prop_attn` (log-size key bias, default True whenever
sizes exist).  A stack cannot be an `nn.scan` (the carry shrinks): it is unrolled natively in csrc/stack.cu.
"""
from __future__ import annotations

import warnings
from typing import Any, Dict, Optional

import numpy as np
import torch

from . import _functional as F
from ._module import AttentionSpec, Module, make_init, merge_param
from .attention import AddPositionEmbedding, Encoder1DBlock, MLPBlock  # noqa: F401  (same classes, re-exported)
from .attention import StackedEncoder1DBlock as _VanillaStack
from .attention import _seed_of

__all__ = ["ToMeMultiHeadDotProductAttention", "ToMeEncoder1DBlock", "AddPositionEmbedding", "StackedEncoder1DBlock"]


class ToMeMultiHeadDotProductAttention(Module):
    """Multi-head dot-product attention with ToMe (tome_attention.py:19-300).

    `__call__` semantics follow the reference: self-attention when `inputs_k`/`inputs_v` are None, the same
    ValueErrors for inconsistent arguments (:94-118), `mask` of shape [batch..., num_heads, q, kv] (or a GroupMask).
    Additive: `size` (fp32 [B, T] token sizes -> proportional attention) and `return_metric` (also return the keys
    reduced over heads, the matching metric of :253)."""

    def __init__(self, num_heads: int, dtype=None, param_dtype="float32", qkv_features: Optional[int] = None,
                 out_features: Optional[int] = None, broadcast_dropout: bool = True, dropout_rate: float = 0.0,
                 deterministic: Optional[bool] = None, precision=None, kernel_init="lecun_normal", bias_init="zeros",
                 use_bias: bool = True, attention_fn=None, decode: bool = False, normalize_qk: bool = False,
                 qkv_dot_general=None, out_dot_general=None, qkv_dot_general_cls=None, out_dot_general_cls=None,
                 tome_r: int = 0, prop_attn: bool = True):
        if decode:
            raise NotImplementedError("decode=True (autoregressive cache, :184-236) is outside the training path")
        if normalize_qk:
            raise NotImplementedError("normalize_qk=True (:166-180) is not implemented")
        if attention_fn is not None:
            raise NotImplementedError("attention_fn is fixed: the fused tcgen05 kernel replaces dot_product_attention (:259-285)")
        self.num_heads, self.qkv_features, self.out_features = num_heads, qkv_features, out_features
        self.broadcast_dropout, self.dropout_rate, self.deterministic = broadcast_dropout, dropout_rate, deterministic
        self.kernel_init, self.bias_init, self.use_bias = kernel_init, bias_init, use_bias
        self.tome_r, self.prop_attn = tome_r, prop_attn

    def _spec(self) -> AttentionSpec:
        return AttentionSpec(self.num_heads, self.qkv_features, self.out_features, self.dropout_rate, self.broadcast_dropout,
                             self.use_bias, False, False, self.kernel_init, self.bias_init, True)

    def _init(self, rng, inputs_q, *a, **k):
        c = inputs_q.shape[-1]
        hd = self.qkv_features or c
        if hd % self.num_heads:
            raise ValueError(f"Memory dimension ({hd}) must be divisible by number of heads ({self.num_heads}).")
        h, d, out = self.num_heads, hd // self.num_heads, self.out_features or c
        ki = make_init(self.kernel_init) if isinstance(self.kernel_init, str) else self.kernel_init
        bi = make_init(self.bias_init) if isinstance(self.bias_init, str) else self.bias_init
        p = {n: {"kernel": ki(rng, (c, h, d), c, hd)} for n in ("query", "key", "value")}
        p["out"] = {"kernel": ki(rng, (h, d, out), hd, out)}
        if self.use_bias:
            for n in ("query", "key", "value"):
                p[n]["bias"] = bi(rng, (h, d))
            p["out"]["bias"] = bi(rng, (out,))
        return p

    def _apply(self, params, inputs_q, inputs_k=None, inputs_v=None, *, inputs_kv=None, mask=None, deterministic=None,
               dropout_rng=None, sow_weights: bool = False, size: Optional[torch.Tensor] = None, return_metric: bool = False):
        if inputs_kv is not None:                                                                      # :94-113
            if inputs_k is not None or inputs_v is not None:
                raise ValueError("If either `inputs_k` or `inputs_v` is not None, `inputs_kv` must be None. If `inputs_kv` is not "
                                 "None, both `inputs_k` and `inputs_v` must be None.")
            inputs_k = inputs_v = inputs_kv
            warnings.warn("The inputs_kv arg will be deprecated soon. Use inputs_k and inputs_v instead.", DeprecationWarning)
        else:
            if inputs_k is None:
                if inputs_v is not None:
                    raise ValueError("`inputs_k` cannot be None if `inputs_v` is not None.")                # :116-121
                inputs_k = inputs_q
            if inputs_v is None:
                inputs_v = inputs_k
        if inputs_k is not inputs_q or inputs_v is not inputs_q:
            raise NotImplementedError("cross-attention: the block path is self-attention only (attention.py:59 passes (x, x))")
        if sow_weights:
            raise NotImplementedError("sow_weights: attention weights are never materialised by the fused kernel")
        if self.dropout_rate > 0.0:                                                                     # :238-247
            if not merge_param("deterministic", self.deterministic, deterministic):
                raise NotImplementedError("attention-weight dropout is not implemented by the fused attention kernel")
        if not inputs_q.is_cuda:
            raise RuntimeError("ToMeMultiHeadDotProductAttention runs on CUDA (sm_100a) only; there is no CPU fallback")
        B, T, C = inputs_q.shape
        x = F._bf16(inputs_q).contiguous()
        gm = F.as_group_mask(mask, B, x.device)
        o, qkv = F.attention(params, self._spec(), x, gm, size if self.prop_attn else None)
        out = F.dense(params["out"], o.reshape(B * T, -1)).view(B, T, -1)                               # :287-299
        if return_metric:
            return out, qkv[:, :, 1].float().sum(dim=2)   # keys summed over heads (:253); cosine matching ignores the scale
        return out


class ToMeEncoder1DBlock(Encoder1DBlock):
    """Transformer encoder layer with a ToMe merge between attention and MLP (tome_attention.py:305-333).

    `apply(variables, inputs, mask=None, train=None, tome_state=ToMeState(), r=...)` -> `(x + y, None)`; the token sizes,
    the groups that follow the merges and the per-block `merge` closures live in `tome_state` (additive argument)."""

    tome = True

    def __init__(self, layer_norm, dropout, self_attention, mlp_block, train: Optional[bool] = None, mask=None,
                 tome_r: int = 0, **kw):
        super().__init__(layer_norm, dropout, self_attention, mlp_block, train, mask, **kw)
        self.tome_r = tome_r

    def _apply(self, params, inputs, mask=None, train=None, r: Optional[int] = None, **kw):
        return super()._apply(params, inputs, mask=mask, train=train, r=self.tome_r if r is None else r, **kw)


class StackedEncoder1DBlock(_VanillaStack):
    """Stacking ToMe encoder layers (tome_attention.py:351-383), unrolled: T shrinks by `tome_r` per block.

    `apply(variables, x, train=False, mask=None)` returns the final tokens [B, T - sum r, C]; `last_size` holds their
    sizes, `unmerge_readouts(idx)` / the engine's readout path recover rows of the original positions."""

    block_cls = ToMeEncoder1DBlock

    def __init__(self, num_blocks: int, encoder_1d_block: Dict[str, Any], tome_r: int = 0, prop_attn: bool = True, **extra):
        super().__init__(num_blocks, encoder_1d_block, prop_attn=prop_attn, **extra)
        self.r = int(tome_r)

    def _block(self, train=None, mask=None):
        cfg = {k: v for k, v in self.encoder_1d_block.items() if k != "_target_"}
        return ToMeEncoder1DBlock(train=train, mask=mask, tome_r=self.r,
                                  attention_dropout=self._extra.get("attention_dropout", "error"), **cfg)
