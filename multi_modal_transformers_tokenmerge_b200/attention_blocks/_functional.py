"""One block's forward as a sequence of C-ABI kernel launches (same kernels and order as csrc/stack.cu, which is
the fast path for whole stacks and for training).  Used by the per-module API (`ToMeMultiHeadDotProductAttention`,
`Encoder1DBlock`, `ToMeEncoder1DBlock`, `MLPBlock`), where the caller owns the loop over blocks.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from .. import _lib as L
from .. import ops
from ._module import AttentionSpec, LayerNormSpec

SUPPORTED_HEAD_DIMS = tuple(range(8, 257, 8))   # 64 runs on the tcgen05 kernels; the others (e.g. the literal 3 x 256) on the generic path


@dataclass
class GroupMask:
    """The block-causal mask as the kernels consume it (instead of octo.py:66-68's [B, H, T, T] booleans):
    gid u8 [B, T] (or [T]), pos i32 [B, T] (or [T]), allow u8 [G, G] with codes 0 masked / 1 visible / 2 causal."""

    gid: torch.Tensor
    pos: torch.Tensor
    allow: torch.Tensor

    def batched(self, batch: int) -> "GroupMask":
        if self.gid.dim() == 2:
            return self
        return GroupMask(self.gid[None].expand(batch, -1).contiguous(), self.pos[None].expand(batch, -1).contiguous(), self.allow)

    @staticmethod
    def from_token_sequence(ts, device="cuda") -> "GroupMask":
        gid, pos = ts.group_ids()
        return GroupMask(torch.as_tensor(gid).to(device), torch.as_tensor(pos).to(device),
                         torch.as_tensor(ts.allow_table()).contiguous().to(device))


def group_mask_from_dense(mask, device="cuda") -> GroupMask:
    """Dense boolean mask ([T,T], [H,T,T] or [B,H,T,T], as built at octo.py:66-68,119) -> GroupMask.  Tokens with the
    same mask row AND column form a group; the mask must be identical over batch and heads and needs <= 32 groups
    (true of every mask token_sequencer.py can generate without a Text set; pass a GroupMask for causal sets)."""
    m = mask.detach().cpu().numpy() if isinstance(mask, torch.Tensor) else np.asarray(mask)
    m = m.astype(bool)
    while m.ndim > 2:
        if not (m == m[:1]).all():
            raise ValueError("dense mask differs across batch/heads: only sequence-structured masks are supported")
        m = m[0]
    T = m.shape[0]
    if m.shape != (T, T):
        raise ValueError(f"self-attention mask must be square, got {m.shape}")
    sig = np.concatenate([m, m.T], axis=1)
    _, first, gid = np.unique(sig, axis=0, return_index=True, return_inverse=True)
    gid = np.asarray(gid).reshape(-1)
    G = len(first)
    if G > 32:
        raise ValueError(f"dense mask has {G} distinct token classes (> 32): pass a GroupMask built from the TokenSequence")
    allow = m[np.ix_(first, first)].astype(np.uint8)
    assert (allow[gid][:, gid] == m).all()
    return GroupMask(torch.as_tensor(gid.astype(np.uint8)).to(device), torch.zeros(T, dtype=torch.int32, device=device),
                     torch.as_tensor(allow).contiguous().to(device))


def as_group_mask(mask, batch: int, device) -> Optional[GroupMask]:
    if mask is None:
        return None
    gm = mask if isinstance(mask, GroupMask) else group_mask_from_dense(mask, device)
    return gm.batched(batch)


@dataclass
class ToMeState:
    """What a ToMe stack has to carry from block to block (the reference has no place for it: tome_attention.py:250-251
    "size should be defined in the model and returned in the end")."""

    size: Optional[torch.Tensor] = None           # f32 [B, T]; None = all ones
    mask: Optional[GroupMask] = None              # groups follow the merges
    merges: List[object] = field(default_factory=list)  # one `merge` closure per block (for unmerge)


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.bfloat16 else t.to(torch.bfloat16)


def _w(t) -> torch.Tensor:  # parameter leaf -> contiguous bf16 [in, out] matrix on the device
    return _bf16(torch.as_tensor(t).cuda()).contiguous()


def _f(t) -> torch.Tensor:
    return torch.as_tensor(t).cuda().float().contiguous()


def layer_norm(p, spec: LayerNormSpec, x: torch.Tensor) -> torch.Tensor:
    y, _, _ = ops.layernorm_fwd(x, _f(p["scale"]), _f(p["bias"]), eps=spec.epsilon, axis=spec.axis)
    return y


def dense(p, x2d: torch.Tensor, *, relu=False, residual=None, dropout=0.0, seed=0, site=0) -> torch.Tensor:
    """flax Dense / DenseGeneral: kernel [in..., out...] flattened to [K, N]; bias [N]."""
    kern = p["kernel"]
    k = x2d.shape[1]
    w = _w(kern).reshape(k, -1)
    bias = _f(p["bias"]).reshape(-1) if "bias" in p else None
    return ops.gemm(x2d, w, m=x2d.shape[0], n=w.shape[1], k=k, b_major=L.TOME_MAJOR_MN, bias=bias, relu=relu,
                    residual=residual, dropout_rate=dropout, dropout_seed=seed, dropout_site=site)


ATTN_DROP_SITE = 0x40000000   # attention-weight dropout of layer l draws from site ATTN_DROP_SITE + l (csrc/stack.cu)


def attention(p, spec: AttentionSpec, x: torch.Tensor, mask: Optional[GroupMask], size: Optional[torch.Tensor],
              dropout_rate: float = 0.0, dropout_seed: int = 0, layer: int = 0):
    """query/key/value DenseGeneral (tome_attention.py:145-164) + dot_product_attention with the group mask and the
    log(size) bias (:259-285).  Returns (o [B,T,H*D] before the `out` projection, packed qkv [B,T,3,H,D])."""
    B, T, C = x.shape
    H = spec.num_heads
    qkv_features = spec.qkv_features or C
    if qkv_features % H:
        raise ValueError(f"Memory dimension ({qkv_features}) must be divisible by number of heads ({H}).")  # :139-142
    D = qkv_features // H
    if D not in SUPPORTED_HEAD_DIMS:
        raise NotImplementedError(f"head_dim {D}: the attention kernels take multiples of 8 up to 256")
    wqkv = torch.cat([_w(p[n]["kernel"]).reshape(C, H * D) for n in ("query", "key", "value")], dim=1).contiguous()
    bqkv = torch.cat([_f(p[n]["bias"]).reshape(-1) for n in ("query", "key", "value")]) if spec.use_bias else None
    qkv = ops.gemm(x.reshape(B * T, C), wqkv, m=B * T, n=3 * H * D, k=C, b_major=L.TOME_MAJOR_MN, bias=bqkv)
    qkv = qkv.view(B, T, 3, H, D)
    kw = {}
    if mask is not None:
        kw = dict(gid=mask.gid, pos=mask.pos, allow=mask.allow)
    o, _ = ops.attention_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], size=size, scale=1.0 / math.sqrt(D),
                             dropout_rate=dropout_rate, dropout_seed=dropout_seed, dropout_site=ATTN_DROP_SITE + layer, **kw)
    return o.view(B, T, H * D), qkv


def tome_merge(qkv: torch.Tensor, x1: torch.Tensor, r: int, state: ToMeState, class_token=False, distill_token=False):
    """The intended ToMe step (tome_attention.py:249-256): metric = keys reduced over heads, read in place from the packed
    qkv buffer; merge the residual stream; sizes and groups follow."""
    from ..tokenizers.token_compression import _Merge

    B, T, _, H, D = qkv.shape
    r = ops.clamp_r(T, r, class_token, distill_token)
    if r <= 0:
        state.merges.append(_Merge(None, T))
        return x1
    nm, ni, _ = ops.sim_argmax(qkv, heads=H, dim=D, batch=B, tokens=T, batch_stride=T * 3 * H * D, token_stride=3 * H * D,
                               head_stride=D, offset_elems=H * D, class_token=class_token, distill_token=distill_token)
    plan = ops.select_topr(nm, ni, T, r, distill_token=distill_token)
    gm = state.mask
    y, size, gid, pos = ops.merge_fwd(plan, x1, state.size, L.TOME_MERGE_WAVG, None if gm is None else gm.gid,
                                      None if gm is None else gm.pos)
    state.size = size
    if gm is not None:
        state.mask = GroupMask(gid, pos, gm.allow)
    state.merges.append(_Merge(plan, T))
    return y
