"""CPU oracle for the ToMe transformer block  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import
this file.  The product package (`multi_modal_transformers_tokenmerge_b200`) never does.

It restates, line by line, the reference's algorithm for the hot path (paths relative to
/root/reference/multi_modal_transformers/):

  * matching / merge  ............ tokenizers/token_compression.py:54-129   (numpy fp32, sequential loops)
  * block / MLP / stack .......... attention_blocks/attention.py:20-119      (torch CPU, autograd gives grads)
  * MHA projections .............. attention_blocks/tome_attention.py:137-164, 259-299
  * hyper-parameters, LN axes .... model_configs/attention_blocks/vanilla_decoder.yaml:1-59
  * mask rules ................... tokenizers/token_sequencer.py:55-183, 199-253, 313-334
  * mask use / readout gather .... models/octo/octo.py:66-68, 116-126
  * action heads + losses ........ action_heads/continuous.py:16-25, action_heads/categorical.py:12-40,
                                   models/octo/octo.py:157-165, 178-190 (l2 / cross-entropy), :253-263, 292-303 (mean);
                                   action_heads/diffusion.py:16-64, 85-143 (schedule, Fourier features, denoiser, loss)

Third-party arithmetic that is NOT under /root/reference and is restated from its published behaviour:
flax ^0.8.2 (`dot_product_attention`, `DenseGeneral`, `LayerNorm(use_fast_variance=True)`, `Dropout`),
jax ^0.4.26 (`argsort` stable, `argmax` first-max, `.at[].add`), see pyproject.toml:27-47.

Parity pinning: the matching/merge functions are checked against golden vectors produced by EXECUTING the
reference's own `token_compression.py` / `token_sequencer.py` / `action_heads/continuous.py` /
`action_heads/categorical.py` / `action_heads/diffusion.py` under a numpy shim of jax + flax (oracle/gen_golden.py -> tests/golden/*.npz) and against the hand-checked vector of SURVEY.md Appendix B.  The block-level pieces that do
not exist in runnable form in the reference (ToMe placement, `unmerge`, proportional `log size` bias -- the
reference's tome_attention.py is a SyntaxError and has no tests) are DEFINED here following the ToMe paper;
for those rows parity is "unpinned by the reference" and this file is the only pin.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------------------------
# 1. bipartite soft matching + merge   (token_compression.py:54-129)
# --------------------------------------------------------------------------------------------------------


@dataclass
class MatchPlan:
    """Everything `bipartite_soft_matching` closes over (token_compression.py:84-88) plus the scores."""

    t: int
    r: int
    distill_token: bool
    scores: Optional[np.ndarray]  # [B,Ta,Tb] fp32 (None when r == 0)
    node_max: Optional[np.ndarray]  # [B,Ta] fp32
    node_idx: Optional[np.ndarray]  # [B,Ta] int32
    edge_idx: Optional[np.ndarray]  # [B,Ta] int32  (full ranking, value desc / index desc)
    unm_idx: Optional[np.ndarray]  # [B,Ta-r] int32
    src_idx: Optional[np.ndarray]  # [B,r] int32
    dst_idx: Optional[np.ndarray]  # [B,r] int32


def clamp_r(t: int, r: int, class_token: bool = False, distill_token: bool = False) -> int:
    """token_compression.py:60-67."""
    protected = int(bool(class_token)) + int(bool(distill_token))
    return max(0, min(int(r), (t - protected) // 2))


def similarity_scores(metric: np.ndarray, class_token=False, distill_token=False) -> np.ndarray:
    """token_compression.py:72-80: L2-normalise (no epsilon), even/odd split, a @ b^T, protect rows/cols."""
    metric = np.asarray(metric, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        metric = metric / np.sqrt(np.sum(metric * metric, axis=-1, keepdims=True, dtype=np.float32))
    a, b = metric[..., ::2, :], metric[..., 1::2, :]
    scores = np.matmul(a, np.swapaxes(b, -1, -2)).astype(np.float32)
    if class_token:
        scores[..., 0, :] = -np.inf
    if distill_token:
        scores[..., :, 0] = -np.inf
    return scores


def plan_from_node(node_max: np.ndarray, node_idx: np.ndarray, t: int, r: int, distill_token: bool = False) -> MatchPlan:
    """token_compression.py:84-88 given the row max / arg max (used when the test follows the GPU's matching
    decisions layer by layer: the ranking and index split are still recomputed here)."""
    return _plan_from_node(None, np.asarray(node_max, np.float32), np.asarray(node_idx, np.int32), t, r, distill_token)


def plan_from_scores(scores: np.ndarray, t: int, r: int, distill_token: bool = False) -> MatchPlan:
    """token_compression.py:82-88 given fp32 scores [B,Ta,Tb]; `r` must already be clamped."""
    scores = np.asarray(scores, dtype=np.float32)
    node_max = scores.max(axis=-1)
    node_idx = scores.argmax(axis=-1).astype(np.int32)  # first maximum (NaN counts as maximum, like XLA)
    return _plan_from_node(scores, node_max, node_idx, t, r, distill_token)


def _plan_from_node(scores, node_max, node_idx, t, r, distill_token):
    # jnp.argsort is stable ascending with NaNs last; [:, ::-1] => value descending, ties by index descending
    edge_idx = np.argsort(node_max, axis=-1, kind="stable")[:, ::-1].astype(np.int32)
    unm_idx = edge_idx[:, r:]
    src_idx = edge_idx[:, :r]
    dst_idx = np.take_along_axis(node_idx, src_idx, axis=-1)
    return MatchPlan(t, r, distill_token, scores, node_max, node_idx, edge_idx, unm_idx, src_idx, dst_idx)


def bipartite_soft_matching(metric, r, class_token=False, distill_token=False, scores_override=None,
                            node_override=None) -> MatchPlan:
    """token_compression.py:54-112.  `r <= 0` returns an identity plan (the reference returns a tuple there,
    which `merge_wavg` cannot call -- SURVEY Appendix C; identity is the documented fix).
    `scores_override` lets a test inject the GPU's fp32 scores so index parity is judged on identical scores."""
    t = int(np.shape(metric)[1])
    r = clamp_r(t, r, class_token, distill_token)
    if r <= 0:
        return MatchPlan(t, 0, distill_token, None, None, None, None, None, None, None)
    if node_override is not None:
        return plan_from_node(node_override[0], node_override[1], t, r, distill_token)
    scores = similarity_scores(metric, class_token, distill_token) if scores_override is None else scores_override
    return plan_from_scores(scores, t, r, distill_token)


def merge(plan: MatchPlan, x: np.ndarray, mode: str = "sum") -> np.ndarray:
    """The closure `merge` of token_compression.py:90-109 (fp32, sequential scatter in rank order)."""
    if plan.r == 0:
        return x
    x = np.asarray(x)
    n, t, c = x.shape
    xe, xo = x[:, ::2, :], x[:, 1::2, :]
    unm = np.take_along_axis(xe, plan.unm_idx[..., None], axis=1)
    src = np.take_along_axis(xe, plan.src_idx[..., None], axis=1)
    dst = np.array(xo, copy=True)
    if mode == "sum":
        bidx = np.arange(n)
        for i in range(plan.r):  # token_compression.py:100-101, one dependent scatter-add per edge
            dst[bidx, plan.dst_idx[:, i], :] += src[:, i, :]
    if plan.distill_token:
        return np.concatenate([unm[:, :1], dst[:, :1], unm[:, 1:], dst[:, 1:]], axis=1)
    return np.concatenate([unm, dst], axis=1)


def merge_wavg(plan: MatchPlan, x: np.ndarray, size: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """token_compression.py:114-129."""
    if size is None:
        size = np.ones_like(x[..., 0, None])
    x = merge(plan, x * size, mode="sum")
    size = merge(plan, size, mode="sum")
    x = x / size
    return x, size


def compute_top_k_tokens(embeddings: np.ndarray, importance_scores: np.ndarray, tokenset_idx, tokenset_k):
    """token_compression.py:15-46 for ONE sequence (embeddings [T, C], scores [T]).  jax.lax.top_k (:31): the k largest
    values in descending order; equal values keep the lower index first (XLA's TopK is stable).  Returns (kept rows, ids)."""
    ids = []
    for k, (start, n) in zip(tokenset_k, tokenset_idx):
        sub = np.asarray(importance_scores[start:start + n])                # dynamic_slice_in_dim (:41)
        order = np.argsort(-sub, kind="stable")[:k]                         # top_k (:31)
        ids.append(order.astype(np.int32) + np.int32(start))                # idx += seq_start_idx (:34)
    ids = np.concatenate(ids, axis=-1)                                      # :44
    return np.take(embeddings, ids, axis=0), ids                            # jnp.take(embeddings, ids, axis=0) (:46)


def row_map(plan: MatchPlan) -> np.ndarray:
    """For every input row t of the layer, the row of the merged output it lands in.  [B,T] int32.
    (Derived from the concatenation order of token_compression.py:103-108.)"""
    assert plan.r > 0
    b = plan.edge_idx.shape[0]
    t = plan.t
    ta, tb = (t + 1) // 2, t // 2
    r = plan.r
    rank = np.empty((b, ta), dtype=np.int32)
    np.put_along_axis(rank, plan.edge_idx, np.broadcast_to(np.arange(ta, dtype=np.int32), (b, ta)), axis=1)
    out = np.empty((b, t), dtype=np.int32)
    n_unm = ta - r

    def pos_unm(i):  # i-th unmerged token
        if plan.distill_token:
            return np.where(i == 0, 0, i + 1)
        return i

    def pos_dst(j):
        if plan.distill_token:
            return np.where(j == 0, 1, n_unm + j)
        return n_unm + j

    node = plan.node_idx
    ev = np.where(rank >= r, pos_unm(rank - r), pos_dst(node))
    out[:, ::2] = ev
    out[:, 1::2] = pos_dst(np.arange(tb, dtype=np.int32))[None, :]
    return out


def unmerge(plan: MatchPlan, xm: np.ndarray) -> np.ndarray:
    """ToMe-paper `unmerge` (NOT in the reference): every original row copies its merged row."""
    if plan.r == 0:
        return xm
    rm = row_map(plan)
    return np.take_along_axis(np.asarray(xm), rm[..., None], axis=1)


# --------------------------------------------------------------------------------------------------------
# 2. token-sequence grammar -> groups, allow table, dense mask, readout indices  (token_sequencer.py)
# --------------------------------------------------------------------------------------------------------

KIND_TDP, KIND_TEXT, KIND_IMAGE, KIND_READOUT = 0, 1, 2, 3
_KIND_BY_NAME = {"TaskDescriptionPrefix": KIND_TDP, "Text": KIND_TEXT, "Image": KIND_IMAGE, "Readout": KIND_READOUT}
_MODALITY = {KIND_TDP: "text", KIND_TEXT: "text", KIND_IMAGE: "images", KIND_READOUT: "readouts"}


def parse_token_sequence(seq: str) -> List[Tuple[int, int, int]]:
    """token_sequencer.py:199-253 (no compression string): list of (kind, num_tokens, timestep)."""
    blocks = re.findall(r"\[(.*?)\]", seq)
    reps = []
    for rep in re.findall(r"(?<=\])(.*?)(?=\[|$)", seq):
        reps.append(1 if rep.strip() == "" else int(re.findall(r"\*(\d+)", rep)[0]))
    out, ts = [], 0
    for block, rep in zip(blocks, reps):
        groups = re.split(r";", block)
        for _ in range(rep):
            for g in groups:
                name = re.search(r"^(.*?)\{", g).group(1).strip()
                out.append((_KIND_BY_NAME[name], int(re.search(r"\d+", g).group()), ts))
            ts += 1
    return out


def allow_rule(qk: int, qt: int, kk: int, kt: int) -> int:
    """Rule table of token_sequencer.py:55-183.  0 = masked, 1 = all-ones, 2 = causal-within-set (Text intra)."""
    same = (qt == kt) and (kk == qk)  # isinstance(key, type(query)) and same timestep -> intra rule
    if qk == KIND_TDP:  # :94-113   (TDP subclasses Text: isinstance(TDP-key, Text-query) is also "intra")
        return 1 if same else 0
    if qk == KIND_TEXT:  # :55-91
        if (qt == kt) and kk in (KIND_TEXT, KIND_TDP):  # isinstance(tokenset, Text) incl. subclass TDP
            return 2
        if kk == KIND_READOUT:
            return 0
        return 1 if kt <= qt else 0
    if qk == KIND_IMAGE:  # :116-148
        if same:
            return 1
        if kk == KIND_READOUT:
            return 0
        return 1 if kt <= qt else 0
    if qk == KIND_READOUT:  # :151-183
        if same:
            return 1
        if kk == KIND_READOUT:
            return 0
        return 1 if kt <= qt else 0
    raise ValueError(qk)


def sequence_groups(seq: str):
    """-> (group_id[T] uint8, pos_in_group[T] int32, allow[G,G] uint8, readout_idx[n] int32)."""
    sets = parse_token_sequence(seq)
    gid, pos = [], []
    for g, (_, n, _) in enumerate(sets):
        gid += [g] * n
        pos += list(range(n))
    G = len(sets)
    allow = np.zeros((G, G), dtype=np.uint8)
    for i, (qk, _, qt) in enumerate(sets):
        for j, (kk, _, kt) in enumerate(sets):
            allow[i, j] = allow_rule(qk, qt, kk, kt)
    ro, cur = [], 0
    for kind, n, _ in sets:  # token_sequencer.py:323-334
        if _MODALITY[kind] == "readouts":
            ro += list(range(cur, cur + n))
        cur += n
    return np.asarray(gid, np.uint8), np.asarray(pos, np.int32), allow, np.asarray(ro, np.int32)


def dense_mask(gid_q, pos_q, gid_k, pos_k, allow) -> np.ndarray:
    """[.., Tq, Tk] bool mask from group ids (token_sequencer.py:313-321 expanded)."""
    a = allow[np.asarray(gid_q)[..., :, None], np.asarray(gid_k)[..., None, :]]
    causal_ok = np.asarray(pos_k)[..., None, :] <= np.asarray(pos_q)[..., :, None]
    return (a == 1) | ((a == 2) & causal_ok)


# --------------------------------------------------------------------------------------------------------
# 3. block / stack in torch (CPU, fp32 or fp64); autograd supplies the reference gradients
# --------------------------------------------------------------------------------------------------------


def _torch():
    import torch

    return torch


def layer_norm(x, scale, bias, eps=1e-6, axis="seq"):
    """flax.linen.LayerNorm as configured in vanilla_decoder.yaml:7-13: reduction_axes=[1] (TOKENS), feature
    axis -1 for scale/bias, use_fast_variance (var = max(0, E[x^2]-E[x]^2)).  axis="feature" is the
    conventional last-axis LayerNorm (opt-in, not what the reference config says)."""
    torch = _torch()
    ax = 1 if axis == "seq" else -1
    mu = x.mean(dim=ax, keepdim=True)
    var = torch.clamp((x * x).mean(dim=ax, keepdim=True) - mu * mu, min=0.0)
    return (x - mu) * torch.rsqrt(var + eps) * scale + bias


def attention(q, k, v, mask=None, bias=None, drop_keep=None):
    """flax dot_product_attention (restated): q,k,v [B,T,H,D]; mask bool broadcastable to [B,H,Tq,Tk];
    bias added to logits before masking; masked logits -> finfo.min; softmax; optional dropout keep-mask
    (already scaled) on the weights; returns [B,T,H,D]."""
    torch = _torch()
    d = q.shape[-1]
    logits = torch.einsum("bqhd,bkhd->bhqk", q / math.sqrt(d), k)
    if bias is not None:
        logits = logits + bias
    if mask is not None:
        logits = torch.where(mask, logits, torch.full_like(logits, torch.finfo(logits.dtype).min))
    w = torch.softmax(logits, dim=-1)
    if drop_keep is not None:
        w = w * drop_keep
    return torch.einsum("bhqk,bkhd->bqhd", w, v)


def dropout_keep_mask(rows: int, cols: int, rate: float, seed: int, site: int):
    """Host restatement of the library's counter-based dropout stream (csrc/common.cuh: drop_stream / DropStream): element
    (row, col) belongs to the stream (row, col // 32); two lowbias32 hashes give the LCG start state and its odd increment;
    element col % 32 is kept iff the top 16 bits of the (col % 32 + 1)-th state are >= round(rate * 65536).
    Returns (keep bool [rows, cols], effective rate).  Used to hand the oracle the exact mask the kernels drew."""
    M = np.uint64(0xFFFFFFFF)

    def lowbias32(x):
        x = x ^ (x >> np.uint64(16)); x = (x * np.uint64(0x21f0aaad)) & M
        x = x ^ (x >> np.uint64(15)); x = (x * np.uint64(0x735a2d97)) & M
        return x ^ (x >> np.uint64(15))

    thresh16 = int(rate * 65536.0 + 0.5)
    thr = np.uint64(thresh16 << 16)
    seed_lo, seed_hi = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    k = (seed_lo ^ ((np.uint64(site) * np.uint64(0x9E3779B9)) & M)) & M
    nch = (cols + 31) // 32
    r = np.arange(rows, dtype=np.uint64)[:, None]
    ch = np.arange(nch, dtype=np.uint64)[None, :]
    x = lowbias32((r * np.uint64(0x9E3779B1) + ch * np.uint64(0x85EBCA77) + k) & M)
    c = lowbias32((((r ^ seed_hi) & M) * np.uint64(0xC2B2AE35) + ch * np.uint64(0x27D4EB2F) + (k ^ np.uint64(0x5bd1e995))) & M) | np.uint64(1)
    keep = np.zeros((rows, nch, 32), bool)
    for e in range(32):
        x = (x * np.uint64(0x915F77F5) + c) & M
        keep[:, :, e] = x >= thr
    return keep.reshape(rows, nch * 32)[:, :cols], thresh16 / 65536.0


def merge_wavg_torch(plan: MatchPlan, x, size):
    """merge_wavg with autograd support; same sequential scatter order as token_compression.py:100-101."""
    torch = _torch()
    if plan.r == 0:
        return x, size

    def _merge(z):
        ze, zo = z[:, ::2, :], z[:, 1::2, :]
        unm_i = torch.as_tensor(plan.unm_idx, dtype=torch.long)[..., None].expand(-1, -1, z.shape[-1])
        src_i = torch.as_tensor(plan.src_idx, dtype=torch.long)[..., None].expand(-1, -1, z.shape[-1])
        unm = torch.gather(ze, 1, unm_i)
        src = torch.gather(ze, 1, src_i)
        dst = zo
        n = z.shape[0]
        bidx = torch.arange(n)
        dsti = torch.as_tensor(plan.dst_idx, dtype=torch.long)
        for i in range(plan.r):
            dst = dst.index_put((bidx, dsti[:, i]), src[:, i, :], accumulate=True)
        if plan.distill_token:
            return torch.cat([unm[:, :1], dst[:, :1], unm[:, 1:], dst[:, 1:]], dim=1)
        return torch.cat([unm, dst], dim=1)

    xs = _merge(x * size)
    s2 = _merge(size)
    return xs / s2, s2


@dataclass
class BlockParams:
    """One block's parameters in the Flax tree layout.  Kernels are [in, out]."""

    ln1_scale: "object"
    ln1_bias: "object"
    wq: "object"  # [C, H*D]  (flax: query/kernel [C,H,D] flattened)
    bq: "object"  # [H*D]
    wk: "object"
    bk: "object"
    wv: "object"
    bv: "object"
    wo: "object"  # [H*D, C]   (flax: out/kernel [H,D,C] flattened)
    bo: "object"  # [C]
    ln2_scale: "object"
    ln2_bias: "object"
    w1: "object"  # [C, Dff]
    b1: "object"
    w2: "object"  # [Dff, C]
    b2: "object"

    def tensors(self):
        return [getattr(self, f) for f in self.__dataclass_fields__]


@dataclass
class LayerTrace:
    plan: MatchPlan
    t_in: int


def _round_st(t, dtype):
    """Round activations to `dtype` with a straight-through gradient: lets the fp32 oracle take the SAME ReLU / merge
    decisions as a kernel pipeline that stores its activations in bf16 (the arithmetic itself stays fp32)."""
    if dtype is None:
        return t
    return t + (t.to(dtype).to(t.dtype) - t).detach()


def tome_block(p: BlockParams, x, size, gid, pos, allow, *, num_heads, r, ln_axis="seq", prop_attn=True,
               class_token=False, distill_token=False, scores_override=None, node_override=None,
               trace: Optional[list] = None, act_dtype=None, relu_gate=None, taps: Optional[dict] = None,
               dropout_rate: float = 0.0, attn_dropout_rate: float = 0.0):
    """ToMeEncoder1DBlock with the ToMe-paper placement (reference shell attention.py:52-69):

        x = x + attn(LN(x), mask(groups), bias = log size)       # dropout 0 (parity mode)
        metric = mean_heads(key);  plan = bipartite_soft_matching(metric, r)
        x, size = merge_wavg(plan, x, size);  groups follow the merge (dst keeps its own group/pos)
        x = x + MLP(LN'(x))

    x [B,T,C] torch; size [B,T,1] torch; gid/pos numpy [B,T]; returns (x, size, gid, pos).
    dropout_rate / attn_dropout_rate > 0 (training, `octo.py:120` always passes train=True) draw torch's own masks at the
    reference's four sites -- attention weights (one [T,T] mask for all batch rows and heads: flax broadcast_dropout), after
    the out projection (attention.py:60), after the ReLU and after dense_out (:34, :37); used by bench.py's CPU arm, never by
    a parity test (the kernels' counter-based masks are compared through dropout_keep_mask instead)."""
    torch = _torch()
    B, T, C = x.shape
    H = num_heads
    drop = (lambda t: torch.nn.functional.dropout(t, dropout_rate)) if dropout_rate > 0 else (lambda t: t)
    rd = lambda t: _round_st(t, act_dtype)  # noqa: E731
    h = rd(layer_norm(x, p.ln1_scale, p.ln1_bias, axis=ln_axis))
    q = rd(h @ p.wq + p.bq).reshape(B, T, H, -1)
    k = rd(h @ p.wk + p.bk).reshape(B, T, H, -1)
    v = rd(h @ p.wv + p.bv).reshape(B, T, H, -1)
    mask = torch.as_tensor(dense_mask(gid, pos, gid, pos, allow))[:, None, :, :]
    bias = torch.log(size[:, None, None, :, 0]) if prop_attn else None
    keep = None
    if attn_dropout_rate > 0:
        keep = (torch.rand(1, 1, T, T) >= attn_dropout_rate).to(x.dtype) / (1.0 - attn_dropout_rate)
    o = rd(attention(q, k, v, mask=mask, bias=bias, drop_keep=keep).reshape(B, T, -1))
    x = rd(x + drop(o @ p.wo + p.bo))
    if taps is not None:   # checker hook: the post-attention residual stream (its gradient is dL/dx1, whose column sum is d bo)
        if x.requires_grad:
            x.retain_grad()
        taps["x1"] = x
    # --- ToMe (intended call site tome_attention.py:249-256): metric = keys reduced over heads
    metric = k.detach().mean(dim=2).to(torch.float32).numpy()
    plan = bipartite_soft_matching(metric, r, class_token, distill_token, scores_override=scores_override,
                                   node_override=node_override)
    if trace is not None:
        trace.append(LayerTrace(plan, T))
    if plan.r > 0:
        x, size = merge_wavg_torch(plan, x, size)
        x = rd(x)
        rm = row_map(plan)
        t2 = T - plan.r
        gid2 = np.zeros((B, t2), dtype=gid.dtype)
        pos2 = np.zeros((B, t2), dtype=pos.dtype)
        # a merged row keeps the group/position of its destination (odd) token; unmerged rows keep their own
        keep = np.ones((B, T), dtype=bool)
        keep[:, ::2] = False
        ta = (T + 1) // 2
        rank = np.empty((B, ta), dtype=np.int32)
        np.put_along_axis(rank, plan.edge_idx, np.broadcast_to(np.arange(ta, dtype=np.int32), (B, ta)), axis=1)
        keep[:, ::2] = rank >= plan.r
        bi, ti = np.nonzero(keep)
        gid2[bi, rm[bi, ti]] = gid[bi, ti]
        pos2[bi, rm[bi, ti]] = pos[bi, ti]
        gid, pos = gid2, pos2
    y = rd(layer_norm(x, p.ln2_scale, p.ln2_bias, axis=ln_axis))
    pre = y @ p.w1 + p.b1
    if relu_gate is None:
        y = rd(drop(torch.relu(pre)))  # attention.py:32-34 (dropout at :34,:37 is the identity in parity mode)
    else:
        # parity protocol for the gradients, as node_override is for the matching: the checker takes the SAME ReLU gate
        # decisions as the implementation under test (bool [B,T',Dff]; a pre-activation within bf16 rounding of zero may
        # fall on either side, and one flipped gate moves a whole row of dW).  The forward value is relu() of the oracle's
        # own pre-activation wherever the gates agree; where they differ |pre| is of the order of the rounding error.
        y = rd(torch.where(torch.as_tensor(relu_gate), pre, torch.zeros_like(pre)))
    y = drop(y @ p.w2 + p.b2)
    return rd(x + y), size, gid, pos


def tome_stack(params: Sequence[BlockParams], pos_embedding, x, gid, pos, allow, *, num_heads, r,
               ln_axis="seq", prop_attn=True, scores_override: Optional[Sequence] = None,
               node_override: Optional[Sequence] = None, trace: Optional[list] = None, act_dtype=None,
               relu_gate: Optional[Sequence] = None, dropout_rate: float = 0.0, attn_dropout_rate: float = 0.0):
    """StackedEncoder1DBlock (attention.py:94-119) unrolled, with shrinking T.  Returns
    (x_final [B,T_L,C], size, origin_row [B,T0] = row of x_final each ORIGINAL token ended up in)."""
    torch = _torch()
    B, T0, C = x.shape
    x = _round_st(x + pos_embedding, act_dtype)
    size = torch.ones(B, T0, 1, dtype=x.dtype)
    gid = np.broadcast_to(gid, (B, T0)).copy()
    pos = np.broadcast_to(pos, (B, T0)).copy()
    origin = np.broadcast_to(np.arange(T0, dtype=np.int32), (B, T0)).copy()
    for li, p in enumerate(params):
        tr: list = []
        so = None if scores_override is None else scores_override[li]
        no = None if node_override is None else node_override[li]
        x, size, gid, pos = tome_block(p, x, size, gid, pos, allow, num_heads=num_heads, r=r, ln_axis=ln_axis,
                                       prop_attn=prop_attn, scores_override=so, node_override=no, trace=tr,
                                       act_dtype=act_dtype, relu_gate=None if relu_gate is None else relu_gate[li],
                                       dropout_rate=dropout_rate, attn_dropout_rate=attn_dropout_rate)
        plan = tr[0].plan
        if plan.r > 0:
            origin = np.take_along_axis(row_map(plan), origin, axis=1)
        if trace is not None:
            trace.append(tr[0])
    return x, size, origin


def attention_weights(q, k, mask=None, bias=None):
    """The softmax weights of `attention` [B,H,Tq,Tk] (flax dot_product_attention_weights, before dropout)."""
    torch = _torch()
    logits = torch.einsum("bqhd,bkhd->bhqk", q / math.sqrt(q.shape[-1]), k)
    if bias is not None:
        logits = logits + bias
    if mask is not None:
        logits = torch.where(mask, logits, torch.full_like(logits, torch.finfo(logits.dtype).min))
    return torch.softmax(logits, dim=-1)


def attention_importance(w, mode: str = "row_mean"):
    """compressed_attention.py:303-306: mean over heads of the mean over the LAST axis (keys) of attn_weights [B,H,Tq,Tk]
    -> [B,T] ("row_mean": 1/T for every token, rows of a softmax sum to one).  "received": the inner mean over queries."""
    inner = w.mean(dim=-1) if mode == "row_mean" else w.mean(dim=-2)          # [B,H,T]
    return inner.mean(dim=-2)


def prune_sets_at(sets, layer: int):
    """Token sets of a pruning stack entering `layer`: sets = [(n at layer 0, c dropped per layer)] -> tokenset_idx
    [(start, n)], tokenset_k (token_sequencer.py:222-238: num_tokens - layer * num_compressed_tokens)."""
    idx, ks, start = [], [], 0
    for n0, c in sets:
        n = n0 - layer * c
        idx.append((start, n))
        ks.append(n - c)
        start += n
    return idx, ks


def prune_block(p: BlockParams, x, gid, pos, allow, *, num_heads, tokenset_idx, tokenset_k, importance="received",
                ln_axis="seq", act_dtype=None, relu_gate=None, ids_override=None, next_groups=None, trace: Optional[dict] = None):
    """The pruning sibling of tome_block (compressed_attention.py:328-358 with :303-308): attention, importance scores from
    its weights, per-set top-k (compute_top_k_tokens) of the post-residual tokens, MLP on the kept ones.  The reference
    prunes the attention output before the out projection and then adds the unpruned residual (shapes cannot agree); the
    same token choice applied to both branches is a gather after the residual add, which is what this does.
    ids_override [B, kept]: take the implementation's keep decisions (scores within rounding of each other may swap).
    next_groups = (gid, pos) of the compression grammar's next layer, else the kept tokens carry their own."""
    torch = _torch()
    B, T, C = x.shape
    H = num_heads
    rd = lambda t: _round_st(t, act_dtype)  # noqa: E731
    h = rd(layer_norm(x, p.ln1_scale, p.ln1_bias, axis=ln_axis))
    q = rd(h @ p.wq + p.bq).reshape(B, T, H, -1)
    k = rd(h @ p.wk + p.bk).reshape(B, T, H, -1)
    v = rd(h @ p.wv + p.bv).reshape(B, T, H, -1)
    mask = torch.as_tensor(dense_mask(gid, pos, gid, pos, allow))[:, None, :, :]
    w = attention_weights(q, k, mask=mask)
    o = rd(torch.einsum("bhqk,bkhd->bqhd", w, v).reshape(B, T, -1))
    x = rd(x + (o @ p.wo + p.bo))
    imp = attention_importance(w.detach(), importance).to(torch.float32).numpy()
    if ids_override is None:
        ids = np.stack([compute_top_k_tokens(np.zeros((T, 1), np.float32), imp[b], tokenset_idx, tokenset_k)[1] for b in range(B)])
    else:
        ids = np.asarray(ids_override)
    if trace is not None:
        trace["importance"], trace["ids"] = imp, ids
    x = torch.gather(x, 1, torch.as_tensor(ids.astype(np.int64))[:, :, None].expand(B, ids.shape[1], C))   # jnp.take (:46)
    if next_groups is not None:
        gid, pos = (np.broadcast_to(a, (B, ids.shape[1])).copy() for a in next_groups)
    else:
        gid, pos = np.take_along_axis(gid, ids, axis=1), np.take_along_axis(pos, ids, axis=1)
    y = rd(layer_norm(x, p.ln2_scale, p.ln2_bias, axis=ln_axis))
    pre = y @ p.w1 + p.b1
    y = rd(torch.relu(pre) if relu_gate is None else torch.where(torch.as_tensor(relu_gate), pre, torch.zeros_like(pre)))
    return rd(x + (y @ p.w2 + p.b2)), gid, pos, ids


def prune_stack(params: Sequence[BlockParams], pos_embedding, x, gid, pos, allow, *, num_heads, sets, importance="received",
                ln_axis="seq", act_dtype=None, relu_gate: Optional[Sequence] = None, ids_override: Optional[Sequence] = None,
                layer_groups: Optional[Sequence] = None, trace: Optional[list] = None):
    """StackedCompressedEncoder1DBlock (compressed_attention.py:377-404) with shrinking T.  layer_groups[l] = (gid, pos) of
    TokenSequence.generate_attention_mask(layer=l) (masks[layer_idx], :399); None: groups are carried with the kept tokens.
    Returns (x_final, origin [B,T0] = row of x_final each original token ended in, -1 when pruned)."""
    B, T0, C = x.shape
    x = _round_st(x + pos_embedding, act_dtype)
    g0, p0 = (gid, pos) if layer_groups is None else layer_groups[0]
    gid = np.broadcast_to(g0, (B, T0)).copy()
    pos = np.broadcast_to(p0, (B, T0)).copy()
    origin = np.broadcast_to(np.arange(T0, dtype=np.int32), (B, T0)).copy()
    for li, p in enumerate(params):
        idx, ks = prune_sets_at(sets, li)
        tr: dict = {}
        nxt = None if layer_groups is None or li + 1 >= len(params) else layer_groups[li + 1]
        T = x.shape[1]
        x, gid2, pos2, ids = prune_block(p, x, gid, pos, allow, num_heads=num_heads, tokenset_idx=idx, tokenset_k=ks,
                                         importance=importance, ln_axis=ln_axis, act_dtype=act_dtype,
                                         relu_gate=None if relu_gate is None else relu_gate[li],
                                         ids_override=None if ids_override is None else ids_override[li], next_groups=nxt, trace=tr)
        if layer_groups is not None and li + 1 >= len(params):
            gid2, pos2 = gid2, pos2          # after the last layer no mask is consumed
        gid, pos = gid2, pos2
        rm = np.full((B, T), -1, np.int32)
        np.put_along_axis(rm, ids, np.broadcast_to(np.arange(ids.shape[1], dtype=np.int32), ids.shape), axis=1)
        origin = np.where(origin >= 0, np.take_along_axis(rm, np.maximum(origin, 0), axis=1), -1)
        if trace is not None:
            trace.append(tr)
    return x, origin


def readout_loss(x_final, origin, readout_idx, target):
    """Stand-in for octo.py:123-124 + :167-174: gather the rows the readout tokens ended up in (unmerge to the
    original positions, then jnp.take), mean squared error against `target` [B, n_readout, C]."""
    torch = _torch()
    rows = torch.as_tensor(origin[:, readout_idx], dtype=torch.long)  # [B, n]
    out = torch.gather(x_final, 1, rows[..., None].expand(-1, -1, x_final.shape[-1]))
    return ((out - target) ** 2).mean(), out


# --------------------------------------------------------------------------------------------------------
# 4. synthetic parameters (initialisers of vanilla_decoder.yaml:25-29,38-42; attention.py:98)
# --------------------------------------------------------------------------------------------------------


def init_block_params(rng: np.random.Generator, C: int, H: int, D: int, Dff: int, dtype=np.float32) -> dict:
    """he_normal kernels (std = sqrt(2/fan_in)), normal(0.01) biases, LN scale 1 / bias 0.  numpy dict."""

    def he(fan_in, shape):
        return (rng.standard_normal(shape) * math.sqrt(2.0 / fan_in)).astype(dtype)

    def nb(shape):
        return (rng.standard_normal(shape) * 0.01).astype(dtype)

    hd = H * D
    return dict(
        ln1_scale=np.ones(C, dtype), ln1_bias=np.zeros(C, dtype),
        wq=he(C, (C, hd)), bq=nb(hd), wk=he(C, (C, hd)), bk=nb(hd), wv=he(C, (C, hd)), bv=nb(hd),
        wo=he(hd, (hd, C)), bo=nb(C),
        ln2_scale=np.ones(C, dtype), ln2_bias=np.zeros(C, dtype),
        w1=he(C, (C, Dff)), b1=nb(Dff), w2=he(Dff, (Dff, C)), b2=nb(C),
    )


def block_params_to_torch(d: dict, dtype=None, requires_grad=False) -> BlockParams:
    torch = _torch()
    kw = {}
    for k, v in d.items():
        t = torch.tensor(np.asarray(v), dtype=dtype or torch.float32)
        t.requires_grad_(requires_grad)
        kw[k] = t
    return BlockParams(**kw)


# --------------------------------------------------------------------------------------------------------
# 5. action heads on the readouts + their losses   (torch CPU: autograd gives the gradients)
# --------------------------------------------------------------------------------------------------------


def continuous_action_head(readouts, kernel, bias, max_action: float):
    """ContinuousActionHead.__call__ (action_heads/continuous.py:16-25): mean over the readout axis, Dense,
    reshape to [B, 1, A], tanh(mean / max_action) * max_action.  readouts [B, n, C]; kernel [C, A]; bias [A]."""
    torch = _torch()
    emb = readouts.mean(dim=-2)                                   # :17
    mean = emb @ kernel + bias                                    # :21  flax Dense
    mean = mean.reshape(mean.shape[0], 1, -1)                     # :22  "batch (seq mean) -> batch seq mean", seq = 1
    return torch.tanh(mean / max_action) * max_action             # :24


def l2_loss(predictions, actions):
    """Octo.compute_l2_loss (octo.py:161-165) -> [B]; the train step takes its mean (:253-263)."""
    torch = _torch()
    predictions = torch.squeeze(predictions)                      # :163
    return ((predictions - actions) ** 2).sum(dim=-1)             # :165


def assign_bins(x: np.ndarray, bounds, num_bins: int) -> np.ndarray:
    """categorical.py:12-22: jnp.digitize against linspace(lo, hi, num_bins + 1) in fp32 -- 1-based for in-range values."""
    bins = np.linspace(np.float32(bounds[0]), np.float32(bounds[1]), num_bins + 1, dtype=np.float32)
    return np.digitize(np.asarray(x, np.float32), bins).astype(np.int32)


def categorical_action_head(readouts, kernel, bias, action_space_dim: int):
    """CategoricalActionHead.__call__ (categorical.py:30-40): "(action timestep)" groups, mean over the timestep axis,
    squeeze, Dense -> logits [B, action, num_bins]."""
    torch = _torch()
    B, n, C = readouts.shape
    emb = readouts.reshape(B, action_space_dim, n // action_space_dim, C)   # :31-35
    emb = torch.squeeze(emb.mean(dim=-2))                                    # :37
    return emb @ kernel + bias                                              # :38


def ce_loss(logits, actions: np.ndarray, max_action: float, num_bins: int):
    """Octo.compute_ce_loss (octo.py:182-190): labels = one_hot(assign_bins(actions)), out-of-range index -> all-zero row
    (jax.nn.one_hot); optax.softmax_cross_entropy = -sum(labels * log_softmax(logits), -1) -> [B, action]."""
    torch = _torch()
    idx = assign_bins(actions, (-max_action, max_action), num_bins)         # :182
    onehot = (idx[..., None] == np.arange(num_bins)).astype(np.float32)      # :183
    labels = torch.as_tensor(onehot, dtype=logits.dtype)
    return -(labels * torch.log_softmax(logits, dim=-1)).sum(dim=-1)         # :187


def cosine_beta_schedule(timesteps: int, s: float = 0.008) -> np.ndarray:
    """diffusion.py:16-26 in fp32 (jnp with x64 disabled)."""
    f = np.float32
    t = np.linspace(f(0), f(timesteps), timesteps + 1, dtype=np.float32) / f(timesteps)
    ac = np.cos((t + f(s)) / f(1 + s) * f(np.pi) * f(0.5)) ** 2
    ac = ac / ac[0]
    return np.clip(f(1) - ac[1:] / ac[:-1], 0, 0.999).astype(np.float32)


def alpha_hats(diffusion_steps: int) -> np.ndarray:
    """DiffusionActionHead.setup (diffusion.py:85-92): cumulative products of 1 - beta."""
    alphas = np.float32(1) - cosine_beta_schedule(diffusion_steps)
    return np.array([np.prod(alphas[: i + 1]) for i in range(diffusion_steps)], np.float32)


def _mlp_block(x, k0, b0, k1, b1):
    """MLPBlock (attention.py:20-39) with train=False: Dense -> relu -> Dense (both Dropouts deterministic)."""
    torch = _torch()
    return torch.relu(x @ k0 + b0) @ k1 + b1


def octo_denoise(noisy_action, time, readout_embedding, p):
    """OctoDenoise.__call__ with FourierFeatures (diffusion.py:29-64), num_blocks = 1.  time int [B, 1];
    p: dict of torch tensors fourier_kernel [F/2, 1], tw1, tb1, tw2, tb2, w1, b1, w2, b2 (kernels [in, out])."""
    torch = _torch()
    x = 2 * math.pi * time.to(torch.float32) @ p["fourier_kernel"].T          # :45
    x = torch.cat([torch.cos(x), torch.sin(x)], dim=-1)                        # :46
    time_embedding = _mlp_block(x, p["tw1"], p["tb1"], p["tw2"], p["tb2"])    # :47
    x = torch.cat([noisy_action, time_embedding, readout_embedding], dim=-1)   # :61
    return _mlp_block(x, p["w1"], p["b1"], p["w2"], p["b2"])                  # :62-63


def denoise_loss(readouts, actions, time, noise, ah: np.ndarray, p):
    """DiffusionActionHead.denoise_loss (diffusion.py:114-143) with the random draws (time int [B, 1], noise [B, A])
    supplied by the caller.  Returns (loss, predictions)."""
    torch = _torch()
    alpha_hat = torch.as_tensor(ah)[time.long()]                               # :131  [B, 1]
    noisy = torch.sqrt(alpha_hat) * actions + torch.sqrt(1 - alpha_hat) * noise   # :132-134
    emb = readouts.mean(dim=-2)                                                # :107
    pred = octo_denoise(noisy, time, emb, p)                                   # :110
    loss = (0.5 * (pred - noise) ** 2).sum(dim=-1).mean()                      # :141-142 optax.l2_loss
    return loss, pred


def diffusion_predict_action(readouts, init, noise, betas: np.ndarray, p, clip: float = 5.0):
    """DiffusionActionHead.predict_action (diffusion.py:146-213) with the random draws supplied: init [B, A] (the Gaussian start,
    :203) and noise [B, A] -- the reference draws the step noise from the SAME per-sample keys at every step (:179), so one
    tensor serves all steps -- or [steps, B, A].  time runs steps - 1 .. 0 (:210); algorithm 2 of arXiv:2006.11239 with
    c1 = 1 / sqrt(alpha_t), c2 = (1 - alpha_t) / sqrt(1 - alpha_hat_t), c3 = sqrt(beta_t) (:183-186); clip to [-5, 5] (:189)."""
    torch = _torch()
    betas = np.asarray(betas, np.float32)
    alphas = np.float32(1) - betas
    ah = np.array([np.prod(alphas[: i + 1]) for i in range(len(betas))], np.float32)
    emb = readouts.mean(dim=-2)
    x = init.clone()
    steps = len(betas)
    for t in range(steps - 1, -1, -1):
        time = torch.full((x.shape[0], 1), t, dtype=torch.int64)
        eps = octo_denoise(x, time, emb, p)
        c1 = float(np.float32(1) / np.sqrt(alphas[t]))
        c2 = float((np.float32(1) - alphas[t]) / np.sqrt(np.float32(1) - ah[t]))
        c3 = float(np.sqrt(betas[t]))
        nz = noise[steps - 1 - t] if noise.dim() == 3 else noise
        x = torch.clamp(c1 * (x - c2 * eps) + c3 * nz, -clip, clip)
    return x


# ------------------------------------------------------------------------------------------------ image patch-embed front end
# multi_modal_transformers/tokenizers/images/image_tokenizer.py.  Pinned by tests/golden/image_tokenizer.npz,
# made by EXECUTING the reference's ImageTokenizer / ResNetV2Block / image_to_patches / encode_patch_position (train=False)
# under the shim, whose Conv / GroupNorm / max_pool / gelu / Embed / Dense leaves restate Flax 0.8.x from memory
# (oracle/jax_shim/flax/linen.py) -- the same "reference control flow, restated leaves" footing as encoder_blocks.npz.
def image_to_patches(image: np.ndarray, patch_size: int, normalize: bool) -> np.ndarray:
    """image_tokenizer.py:35-71.  image [H, W, C] (square, divisible by the patch) -> [(H/p)*(W/p), p, p, C], patches in
    row-major (h, w) order; normalize: 2 * (x / 255) - 1."""
    h, w, c = image.shape
    assert h == w and h % patch_size == 0
    n = h // patch_size
    pt = image.reshape(n, patch_size, n, patch_size, c).transpose(0, 2, 1, 3, 4).reshape(n * n, patch_size, patch_size, c)
    pt = pt.astype(np.float32)
    if normalize:
        pt = (np.float32(2) * (pt / np.float32(255.0))) - np.float32(1.0)                                   # :67
    return pt


def patch_position_tokens(image_size: int, patch_size: int, num_tokens: int):
    """encode_patch_position with train=False (image_tokenizer.py:74-140): for patch k the pixel interval of index
    k % patches_per_dim ("row", :94 -- the fastest-varying index of the (h w) patch order, i.e. the horizontal one) and of
    k // patches_per_dim ("col", :95) is normalised by the image size, scaled to num_tokens - 1, floored (:100, fp32) and the
    midpoint of the quantised interval taken with a floor division (:112-113).  Returns int32 (row_tokens, col_tokens) [n]."""
    ppd = image_size // patch_size
    f = np.float32
    edges = np.arange(0, image_size + patch_size, patch_size)
    q = np.floor((edges.astype(f) / f(image_size)) * f(num_tokens - 1)).astype(f)
    mid = np.floor_divide(q[:-1] + q[1:], f(2)).astype(np.int32)          # per interval
    k = np.arange(ppd * ppd)
    return mid[k % ppd].astype(np.int32), mid[k // ppd].astype(np.int32)


def _conv2d_nhwc(x, kernel, bias, stride, same):
    """flax.linen.Conv (NHWC, kernel [kh, kw, in, out]); SAME pads (k - 1) // 2 low, the rest high (stride 1)."""
    kh, kw = kernel.shape[:2]
    if same:
        x = np.pad(x, ((0, 0), ((kh - 1) // 2, kh - 1 - (kh - 1) // 2), ((kw - 1) // 2, kw - 1 - (kw - 1) // 2), (0, 0)))
    H, W = x.shape[1:3]
    oh, ow = (H - kh) // stride + 1, (W - kw) // stride + 1
    cols = np.empty((x.shape[0], oh, ow, kh * kw * x.shape[3]), np.float32)
    for dy in range(kh):
        for dx in range(kw):
            t = dy * kw + dx
            cols[..., t * x.shape[3]:(t + 1) * x.shape[3]] = x[:, dy:dy + stride * oh:stride, dx:dx + stride * ow:stride, :]
    return (cols.reshape(-1, cols.shape[-1]) @ kernel.reshape(-1, kernel.shape[3]).astype(np.float32)).reshape(
        x.shape[0], oh, ow, kernel.shape[3]) + bias.astype(np.float32)


def gelu_tanh(x):
    """flax.linen.gelu (approximate=True)."""
    x = x.astype(np.float32)
    c = np.float32(np.sqrt(2.0 / np.pi))
    return (np.float32(0.5) * x * (np.float32(1) + np.tanh(c * (x + np.float32(0.044715) * x * x * x)))).astype(np.float32)


def group_norm_flax(x, scale, bias, groups: int, eps: float):
    """flax.linen.GroupNorm with default reduction axes on x [B, ..., C]: per (batch row, group) statistics over EVERY other
    axis (images, patches, H, W and the group's channels), fast variance, per-channel scale / bias."""
    B, C = x.shape[0], x.shape[-1]
    g = x.reshape(B, -1, groups, C // groups).astype(np.float32)
    mu = g.mean(axis=(1, 3), keepdims=True, dtype=np.float32)
    var = np.maximum((g * g).mean(axis=(1, 3), keepdims=True, dtype=np.float32) - mu * mu, 0)
    y = ((g - mu) * (np.float32(1) / np.sqrt(var + np.float32(eps))).astype(np.float32)).reshape(x.shape)
    return (y * scale.astype(np.float32) + bias.astype(np.float32)).astype(np.float32)


def image_tokenizer_fwd(p: dict, image: np.ndarray, *, patch_size: int, position_interval: int, num_groups: int,
                        conv_stride: int = 2, pool_window: int = 3, gn_eps: float = 1e-6, normalize: bool = True) -> np.ndarray:
    """ImageTokenizer.__call__ with train=False (image_tokenizer.py:216-309) around ResNetV2Block.__call__ (:148-190):
    patches -> Conv (VALID, stride 2) -> max_pool (3x3, stride 1, VALID) -> num_blocks x [GroupNorm -> gelu -> Conv 3x3 SAME]
    -> + residual (the pooled tensor) -> flatten -> Dense -> + row / column position embeddings.
    image [B, N, H, W, C] (pixel values 0..255); p: conv0_kernel [k, k, C, F], conv0_bias, blocks = [(gn_scale, gn_bias,
    conv_kernel [3, 3, F, F], conv_bias), ...], dense_kernel [o*o*F, E], dense_bias, row_embedding / col_embedding
    [position_interval, E].  Returns [B, N, n_patches, E] fp32."""
    B, N, H, W, C = image.shape
    pt = np.stack([np.stack([image_to_patches(image[b, i], patch_size, normalize) for i in range(N)]) for b in range(B)])
    n = pt.shape[2]
    x = _conv2d_nhwc(pt.reshape((-1,) + pt.shape[3:]), p["conv0_kernel"], p["conv0_bias"], conv_stride, same=False)
    o1 = x.shape[1]
    o2 = o1 - pool_window + 1
    pooled = np.full((x.shape[0], o2, o2, x.shape[3]), -np.inf, np.float32)
    for dy in range(pool_window):
        for dx in range(pool_window):
            pooled = np.maximum(pooled, x[:, dy:dy + o2, dx:dx + o2, :])
    x = pooled
    for (gs, gb, ck, cb) in p["blocks"]:
        h = group_norm_flax(x.reshape(B, N, n, o2, o2, -1), gs, gb, num_groups, gn_eps).reshape(x.shape)
        x = _conv2d_nhwc(gelu_tanh(h), ck, cb, 1, same=True)
    x = x + pooled                                                             # :170 (shapes agree: no projection of the residual)
    tok = x.reshape(B, N, n, -1) @ p["dense_kernel"].astype(np.float32) + p["dense_bias"].astype(np.float32)
    rt, ct = patch_position_tokens(H, patch_size, position_interval)
    return (tok + p["row_embedding"].astype(np.float32)[rt] + p["col_embedding"].astype(np.float32)[ct]).astype(np.float32)


def image_tokenizer_params_from_flax(tree: dict, num_blocks: int) -> dict:
    """Flax parameter tree of ImageTokenizer (setup-style children named by attribute / explicit `name`, ResNetV2Block's
    compact children auto-named Conv_i / GroupNorm_i / Dense_0) -> the dict image_tokenizer_fwd takes."""
    ef = tree["embedding_function"]
    g = lambda a: np.asarray(a, np.float32)  # noqa: E731
    return dict(conv0_kernel=g(ef["Conv_0"]["kernel"]), conv0_bias=g(ef["Conv_0"]["bias"]),
                blocks=[(g(ef[f"GroupNorm_{i}"]["scale"]), g(ef[f"GroupNorm_{i}"]["bias"]), g(ef[f"Conv_{i + 1}"]["kernel"]),
                         g(ef[f"Conv_{i + 1}"]["bias"])) for i in range(num_blocks)],
                dense_kernel=g(ef["Dense_0"]["kernel"]), dense_bias=g(ef["Dense_0"]["bias"]),
                row_embedding=g(tree["image_row_position_embedding"]["embedding"]),
                col_embedding=g(tree["image_col_position_embedding"]["embedding"]))


# ------------------------------------------------------------------------------------------------ attention pooling
def attention_pooling(x: np.ndarray, p: dict, num_heads: int, ln_axis: int = 1, eps: float = 1e-6) -> np.ndarray:
    """MultiHeadAttentionPooling.__call__ in evaluation mode (attention_blocks/attention.py:122-150): a learnt
    query [1, 1, E] tiled over the batch (:139-144) attends over the tokens x [B, n, E] (flax MultiHeadDotProductAttention: per-head
    Dense projections, q / sqrt(D), softmax over the n keys, out projection; :147), then LayerNorm (:148) -> MLPBlock (:149) on
    the pooled token, returned with the residual (:151).  ln_axis: the configured reduction axis (diffusion.yaml:21 says [1] -- the
    single pooled token, whose normalised value is 0, so LayerNorm returns its bias; -1 = features).
    p: Flax tree {learnt_q_input, MultiHeadDotProductAttention_0: {query, key, value, out}, LayerNorm_0, MLPBlock_0: {Dense_0, Dense_1}}.
    Pinned by tests/golden/attention_pooling.npz (the reference's own module executed under the shim).  Returns [B, 1, E]."""
    f = np.float32
    x = np.asarray(x, f)
    B, n, E = x.shape
    a = p["MultiHeadDotProductAttention_0"]
    g = lambda t: np.asarray(t, f)  # noqa: E731
    q = np.einsum("btc,chd->bthd", np.broadcast_to(g(p["learnt_q_input"]), (B, 1, E)), g(a["query"]["kernel"])) + g(a["query"]["bias"])
    k = np.einsum("btc,chd->bthd", x, g(a["key"]["kernel"])) + g(a["key"]["bias"])
    v = np.einsum("btc,chd->bthd", x, g(a["value"]["kernel"])) + g(a["value"]["bias"])
    logits = np.einsum("bqhd,bkhd->bhqk", q / np.sqrt(f(q.shape[-1])), k)
    w = np.exp(logits - logits.max(-1, keepdims=True))
    w = (w / w.sum(-1, keepdims=True)).astype(f)
    o = np.einsum("bhqk,bkhd->bqhd", w, v)
    x1 = (np.einsum("bthd,hdc->btc", o, g(a["out"]["kernel"])) + g(a["out"]["bias"])).astype(f)       # [B, 1, E]
    ax = 1 if ln_axis == 1 else 2
    mu = x1.mean(axis=ax, keepdims=True, dtype=f)
    var = np.maximum((x1 * x1).mean(axis=ax, keepdims=True, dtype=f) - mu * mu, 0)
    y = ((x1 - mu) / np.sqrt(var + f(eps))).astype(f) * g(p["LayerNorm_0"]["scale"]) + g(p["LayerNorm_0"]["bias"])
    m = p["MLPBlock_0"]
    y = np.maximum(y @ g(m["Dense_0"]["kernel"]) + g(m["Dense_0"]["bias"]), 0)
    y = y @ g(m["Dense_1"]["kernel"]) + g(m["Dense_1"]["bias"])
    return (x1 + y).astype(f)
