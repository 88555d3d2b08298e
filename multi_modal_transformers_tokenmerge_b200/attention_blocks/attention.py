"""The reference's block modules with their names, attributes and call signatures, running on the sm_100a kernels.

Mirror of multi_modal_transformers/attention_blocks/attention.py:
    MLPBlock(dense, activation, norm, dense_out)(inputs, train=False)                         :20-39
    Encoder1DBlock(layer_norm, dropout, self_attention, mlp_block, train, mask)(inputs, mask, train) -> (y, None)   :41-69
    AddPositionEmbedding(posemb_init)(inputs)                                                 :71-85
    StackedEncoder1DBlock(num_blocks, encoder_1d_block)(x, train=False, mask=None)            :87-119

Attributes are the config nodes of model_configs/attention_blocks/vanilla_decoder.yaml (plain dicts, as PyYAML loads
them; hydra's DictConfig behaves the same for the keys read here).  Modules are functional like Flax's:
`variables = m.init(rng, x)`, `y = m.apply(variables, x, ...)`, parameter names as Flax would assign them
(SURVEY.md A.5).  Inputs and outputs are CUDA torch tensors; activations are bf16, parameters fp32.

Differences from the reference that a caller can observe (all listed in DESIGN.md):
  * dropout: hidden dropout (yaml:17,50) and attention-weight dropout (`self_attention.dropout_rate`, yaml:23, with
    flax's default broadcast_dropout=True: one [T,T] mask for all batch rows and heads) are implemented, regenerated in
    backward from (seed, site, element) and never stored -- but the random stream is a counter-based LCG, not Flax's
    threefry; `attention_dropout="ignore"` switches the attention-weight dropout off; broadcast_dropout=False raises;
  * `mask` may be the dense boolean array of octo.py:66-68 (converted to a group table on the host) or, on the fast
    path, a `GroupMask` / `TokenSequence`.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np
import torch

from .. import ops
from . import _functional as F
from ._module import (AttentionSpec, DenseSpec, DropoutSpec, LayerNormSpec, Module, as_rng, call, instantiate, make_init,
                      merge_param)


def _seed_of(dropout_rng) -> int:
    if dropout_rng is None:
        return 0
    return int(np.asarray(dropout_rng).ravel()[-1])


def _dense_init(rng, spec: DenseSpec, fan_in: int, shape_k, shape_b):
    p = {"kernel": make_init(spec.kernel_init)(rng, shape_k, fan_in, int(np.prod(shape_b)))}
    if spec.use_bias:
        p["bias"] = make_init(spec.bias_init)(rng, shape_b)
    return p


class MLPBlock(Module):
    """Transformer MLP / feed-forward block (attention.py:20-39): Dense -> activation -> Dropout -> Dense -> Dropout."""

    def __init__(self, dense: Dict[str, Any], activation: Dict[str, Any], norm: Dict[str, Any], dense_out: Dict[str, Any]):
        self.dense, self.activation, self.norm, self.dense_out = dense, activation, norm, dense_out

    def _specs(self):
        d, do = instantiate(self.dense), instantiate(self.dense_out)
        act, drop = call(self.activation), instantiate(self.norm)
        if not isinstance(drop, DropoutSpec):
            raise ValueError("MLPBlock.norm must be a flax.linen.Dropout node (attention.py:34,37)")
        return d, act, drop, do

    def _init(self, rng, inputs, train=False):
        d, _, _, do = self._specs()
        c = inputs.shape[-1]
        return {"Dense_0": _dense_init(rng, d, c, (c, d.features), (d.features,)),
                "Dense_1": _dense_init(rng, do, d.features, (d.features, do.features), (do.features,))}

    def _apply(self, params, inputs, train=False, residual=None, dropout_rng=None, site=0):
        d, act, drop, do = self._specs()
        shp = inputs.shape
        x2 = F._bf16(inputs).reshape(-1, shp[-1]).contiguous()
        rate = drop.rate if train else 0.0
        seed = _seed_of(dropout_rng)
        h = F.dense(params["Dense_0"], x2, relu=(act == "relu"), dropout=rate, seed=seed, site=site + 1)       # :32-34
        res2 = None if residual is None else residual.reshape(-1, do.features)
        y = F.dense(params["Dense_1"], h, residual=res2, dropout=rate, seed=seed, site=site + 2)                # :36-37
        return y.view(*shp[:-1], do.features)


class MultiHeadAttentionPooling(Module):
    """Multihead attention pooling (attention.py:122-150): a learnt query attends over the tokens, then LayerNorm -> MLPBlock on
    the pooled token with the residual.  Forward (the reference's heads keep its call commented out: continuous.py:18-19,
    diffusion.py:99-100).  Parameters, Flax names: learnt_q_input [1, 1, E], MultiHeadDotProductAttention_0/{query, key, value,
    out}, LayerNorm_0, MLPBlock_0.  Key / value / out projections, LayerNorm and the MLP are the library's GEMM / LayerNorm
    kernels; the one-query attention itself is csrc/attn_pool.cu."""

    def __init__(self, query_map_input: Dict[str, Any], dot_product_attention: Dict[str, Any], layer_norm: Dict[str, Any],
                 mlp_block: Dict[str, Any]):
        self.query_map_input, self.dot_product_attention = query_map_input, dot_product_attention
        self.layer_norm, self.mlp_block = layer_norm, mlp_block

    def _specs(self):
        at, ln = instantiate(self.dot_product_attention), instantiate(self.layer_norm)
        assert isinstance(at, AttentionSpec) and isinstance(ln, LayerNormSpec)
        if at.tome:
            raise ValueError("MultiHeadAttentionPooling.dot_product_attention must be a flax attention node")
        return at, ln, MLPBlock(**{k: v for k, v in self.mlp_block.items() if k != "_target_"})

    def _init(self, rng, x, train=False):
        at, ln, mlp = self._specs()
        E = x.shape[-1]
        H = at.num_heads
        D = (at.qkv_features or E) // H
        from ._module import _init_name
        qinit = make_init(_init_name(self.query_map_input.get("kernel_init"), "lecun_normal"))
        proj = lambda: {"kernel": make_init(at.kernel_init)(rng, (E, H, D), E, H * D), "bias": make_init(at.bias_init)(rng, (H, D))}  # noqa: E731
        attn = {"query": proj(), "key": proj(), "value": proj(),
                "out": {"kernel": make_init(at.kernel_init)(rng, (H, D, E), H * D, E), "bias": make_init(at.bias_init)(rng, (E,))}}
        return {"learnt_q_input": qinit(rng, (1, 1, E), E, E), "MultiHeadDotProductAttention_0": attn,
                "LayerNorm_0": {"scale": np.ones(E, np.float32), "bias": np.zeros(E, np.float32)},
                "MLPBlock_0": mlp._init(rng, np.zeros((1, 1, E), np.float32))}

    def _apply(self, params, x, train=False, dropout_rng=None):
        import ctypes as C_

        from .. import _lib as L_
        at, ln, mlp = self._specs()
        if not x.is_cuda:
            raise RuntimeError("MultiHeadAttentionPooling runs on CUDA (sm_100a) only: got a CPU tensor.  There is no CPU fallback.")
        if x.dim() != 3:
            raise ValueError("x must be [batch, sequence, embedding] (attention.py:135)")
        if train and at.dropout_rate > 0.0:
            raise NotImplementedError("attention-weight dropout inside the pooling attention is not implemented")
        B, n, E = x.shape
        H = at.num_heads
        hd = at.qkv_features or E
        if hd % H:
            raise ValueError(f"Memory dimension ({hd}) must be divisible by number of heads ({H}).")
        D = hd // H
        a = params["MultiHeadDotProductAttention_0"]
        dev = x.device
        # key | value projections of the tokens in one GEMM (flax DenseGeneral [E, H, D] kernels side by side)
        wkv = torch.cat([F._w(a[k_]["kernel"]).reshape(E, hd) for k_ in ("key", "value")], dim=1).contiguous()
        bkv = torch.cat([F._f(a[k_]["bias"]).reshape(-1) for k_ in ("key", "value")])
        kv = ops.gemm(F._bf16(x).reshape(B * n, E).contiguous(), wkv, m=B * n, n=2 * hd, k=E, b_major=L_.TOME_MAJOR_MN, bias=bkv)
        learnt = F._f(params["learnt_q_input"]).reshape(E).contiguous()
        wq = F._f(a["query"]["kernel"]).reshape(E, hd).contiguous()
        bq = F._f(a["query"]["bias"]).reshape(hd).contiguous()
        q_scratch = torch.empty(hd, dtype=torch.float32, device=dev)
        o = torch.empty(B, hd, dtype=torch.bfloat16, device=dev)
        L_.check(L_.lib().tome_attention_pool_fwd(B, n, H, D, E, C_.c_void_p(learnt.data_ptr()), C_.c_void_p(wq.data_ptr()),
                                                C_.c_void_p(bq.data_ptr()), C_.c_void_p(kv.data_ptr()), 2 * hd, hd,
                                                C_.c_void_p(q_scratch.data_ptr()), C_.c_void_p(o.data_ptr()),
                                                C_.c_void_p(torch.cuda.current_stream().cuda_stream)))
        x1 = F.dense(a["out"], o)                                                        # [B, E]            (:147)
        y = F.layer_norm(params["LayerNorm_0"], ln, x1.view(B, 1, E))                    # (:148)
        y = mlp._apply(params["MLPBlock_0"], y, train=train, residual=x1.view(B, 1, E), dropout_rng=dropout_rng)   # (:149-151)
        return y


class Encoder1DBlock(Module):
    """Transformer encoder layer (attention.py:41-69); returns `(x + y, None)` like the scanned reference block."""

    tome = False

    def __init__(self, layer_norm, dropout, self_attention, mlp_block, train: Optional[bool] = None, mask=None,
                 attention_dropout: str = "error"):
        self.layer_norm, self.dropout, self.self_attention, self.mlp_block = layer_norm, dropout, self_attention, mlp_block
        self.train, self.mask = train, mask
        self.attention_dropout = attention_dropout

    # -- config ------------------------------------------------------------------------------------------------
    def _specs(self):
        ln, dr, at = instantiate(self.layer_norm), instantiate(self.dropout), instantiate(self.self_attention)
        assert isinstance(ln, LayerNormSpec) and isinstance(dr, DropoutSpec) and isinstance(at, AttentionSpec)
        mlp_cfg = {k: v for k, v in self.mlp_block.items() if k != "_target_"}
        return ln, dr, at, MLPBlock(**mlp_cfg)

    @property
    def _attn_name(self):
        t = self.self_attention["_target_"].rsplit(".", 1)[-1]
        return f"{t}_0"

    def _init(self, rng, inputs, mask=None, train=None):
        ln, _, at, mlp = self._specs()
        c = inputs.shape[-1]
        hd = at.qkv_features or c
        h, d = at.num_heads, hd // at.num_heads
        ki, bi = make_init(at.kernel_init), make_init(at.bias_init)
        attn = {}
        for n in ("query", "key", "value"):                                   # DenseGeneral(features=(H, D)) :145-164
            attn[n] = {"kernel": ki(rng, (c, h, d), c, hd)}
            if at.use_bias:
                attn[n]["bias"] = bi(rng, (h, d))
        attn["out"] = {"kernel": ki(rng, (h, d, at.out_features or c), hd, at.out_features or c)}   # :287-299
        if at.use_bias:
            attn["out"]["bias"] = bi(rng, (at.out_features or c,))
        one = lambda: {"scale": np.ones(c, np.float32), "bias": np.zeros(c, np.float32)}  # noqa: E731
        return {"LayerNorm_0": one(), self._attn_name: attn, "LayerNorm_1": one(), "MLPBlock_0": mlp._init(rng, inputs)}

    # -- forward -----------------------------------------------------------------------------------------------
    def _apply(self, params, inputs, mask=None, train=None, dropout_rng=None, tome_state: Optional[F.ToMeState] = None,
               r: int = 0, prop_attn: bool = True, site: int = 0):
        train = merge_param("train", self.train, train)
        mask = self.mask if mask is None else mask
        ln, dr, at, mlp = self._specs()
        if not inputs.is_cuda:
            raise RuntimeError("Encoder1DBlock runs on CUDA (sm_100a) only; there is no CPU fallback")
        if train and at.dropout_rate > 0.0 and not at.broadcast_dropout:
            raise NotImplementedError("attention-weight dropout is implemented for flax's default broadcast_dropout=True (one "
                                      "mask for all batch rows and heads); broadcast_dropout=False is not")
        attn_rate = at.dropout_rate if (train and self.attention_dropout != "ignore") else 0.0
        B, T, C = inputs.shape
        x_in = F._bf16(inputs).contiguous()
        st = tome_state if tome_state is not None else F.ToMeState()
        if st.mask is None and mask is not None:
            st.mask = F.as_group_mask(mask, B, inputs.device)
        seed = _seed_of(dropout_rng)
        rate = dr.rate if train else 0.0
        x = F.layer_norm(params["LayerNorm_0"], ln, x_in)                                              # :58
        o, qkv = F.attention(params[self._attn_name], at, x, st.mask, st.size if prop_attn else None,
                             dropout_rate=attn_rate, dropout_seed=seed, layer=site // 3)                # :59
        x1 = F.dense(params[self._attn_name]["out"], o.reshape(B * T, -1), residual=x_in.reshape(B * T, C), dropout=rate,
                     seed=seed, site=site).view(B, T, C)                                               # :60-63
        if self.tome and r > 0:
            x1 = F.tome_merge(qkv, x1, r, st)                          # tome_attention.py:249-256 (intent), SURVEY A.7
        y = F.layer_norm(params["LayerNorm_1"], ln, x1)                                                # :66
        out = mlp._apply(params["MLPBlock_0"], y, train, residual=x1, dropout_rng=dropout_rng, site=site)  # :67-69
        return out, None


class AddPositionEmbedding(Module):
    """Adds learned positional embeddings to the inputs (attention.py:71-85)."""

    def __init__(self, posemb_init=None, name: Optional[str] = None):
        self.posemb_init, self.name = posemb_init, name

    def _init(self, rng, inputs):
        assert inputs.dim() == 3, "Number of dimensions should be 3, but it is: %d" % inputs.dim()
        init = self.posemb_init or (lambda g, shape: (g.standard_normal(shape) * 0.02).astype(np.float32))
        return {"pos_embedding": init(rng, (1, inputs.shape[1], inputs.shape[2]))}

    def _apply(self, params, inputs):
        assert inputs.dim() == 3, "Number of dimensions should be 3, but it is: %d" % inputs.dim()
        import ctypes as C
        from .. import _lib as L
        b, t, c = inputs.shape
        pe = torch.as_tensor(params["pos_embedding"]).cuda().float().reshape(t, c).contiguous()
        x = inputs.contiguous()
        y = torch.empty(b, t, c, dtype=torch.bfloat16, device=x.device)
        L.check(L.lib().tome_add_pos_embedding(b, t, c, x.data_ptr(), L.TOME_BF16 if x.dtype == torch.bfloat16 else L.TOME_F32,
                                               pe.data_ptr(), y.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return y


class StackedEncoder1DBlock(Module):
    """Stacking Transformer encoder layers (attention.py:87-119).  The reference scans one block `num_blocks` times
    with parameters stacked on axis 0; the parameter tree here has the same shape (leading [num_blocks] axis under
    "ScanEncoder1DBlock_0").  Execution goes through the native stack executor (csrc/stack.cu) in one call."""

    block_cls = Encoder1DBlock
    r = 0

    def __init__(self, num_blocks: int, encoder_1d_block: Dict[str, Any], **extra):
        self.num_blocks, self.encoder_1d_block = int(num_blocks), encoder_1d_block
        self._extra = extra
        self._engine = None
        self._engine_key = None
        self._loaded_params = None

    def _block(self, train=None, mask=None):
        cfg = {k: v for k, v in self.encoder_1d_block.items() if k != "_target_"}
        return self.block_cls(train=train, mask=mask, attention_dropout=self._extra.get("attention_dropout", "error"), **cfg)

    def _init(self, rng, x, train=False, mask=None):
        blk = self._block()
        per = [blk._init(rng, x) for _ in range(self.num_blocks)]
        stacked = _tree_stack(per)
        pe = AddPositionEmbedding()._init(rng, x)
        return {"posembed_input": pe, "ScanEncoder1DBlock_0": stacked}

    # engine plumbing --------------------------------------------------------------------------------------------
    def engine_for(self, params, x_shape, mask, train: bool, r: int = 0, prop_attn: bool = True, n_readout: int = 0,
                   readout_idx=None, dropout_seed: int = 0):
        """The native executor configured for this stack, with `params` loaded (cached per shape)."""
        from ..engine import StackConfig, ToMeStackEngine

        blk = self._block()
        ln, dr, at, mlp = blk._specs()
        d, act, mdrop, do = mlp._specs()
        B, T, C = x_shape
        hd = at.qkv_features or C
        if abs(mdrop.rate - dr.rate) > 1e-12:
            raise NotImplementedError("the native stack uses one hidden dropout rate (yaml:17 and :50 agree in the reference)")
        if train and at.dropout_rate > 0.0 and not at.broadcast_dropout:
            raise NotImplementedError("attention-weight dropout needs broadcast_dropout=True; see Encoder1DBlock")
        attn_rate = at.dropout_rate if (train and self._extra.get("attention_dropout", "error") != "ignore") else 0.0
        gm = None if mask is None else (mask if isinstance(mask, F.GroupMask) else F.group_mask_from_dense(mask))
        if gm is not None and gm.gid.dim() != 1:
            raise ValueError("the stack takes one group-id vector [T] (every batch row starts from the same sequence)")
        # the dropout seed is NOT part of the key: a caller that passes a fresh dropout rng per step (as the reference does)
        # must not tear down the engine (workspace, parameters, gradients) on every apply; it is updated in place below
        key = (B, T, C, r, train, prop_attn, n_readout, attn_rate, None if gm is None else gm.allow.shape[0])
        if self._engine is None or self._engine_key != key:
            cfg = StackConfig(batch=B, tokens=T, channels=C, heads=at.num_heads, head_dim=hd // at.num_heads, mlp_dim=d.features,
                              layers=self.num_blocks, r=r, ln_axis=ln.axis, ln_eps=ln.epsilon, prop_attn=prop_attn,
                              num_groups=0 if gm is None else int(gm.allow.shape[0]), n_readout=n_readout,
                              dropout_rate=dr.rate if train else 0.0, dropout_seed=dropout_seed, attn_dropout_rate=attn_rate)
            self._engine = ToMeStackEngine(cfg, gid=None if gm is None else gm.gid.cpu().numpy(),
                                           pos=None if gm is None else gm.pos.cpu().numpy(),
                                           allow=None if gm is None else gm.allow.cpu().numpy(), readout_idx=readout_idx,
                                           training=train)
            self._engine_key = key
            self._loaded_params = None
        self._engine.set_dropout_seed(dropout_seed)
        if self._loaded_params is not params:   # same tree object as last time: the device copy is current
            self._engine.load_params(np.asarray(_np(params["posembed_input"]["pos_embedding"])).reshape(T, C),
                                     flax_tree_to_layers(params["ScanEncoder1DBlock_0"], blk._attn_name, self.num_blocks))
            self._loaded_params = params
        return self._engine

    def _apply(self, params, x, train=False, mask=None, dropout_rng=None):
        if not x.is_cuda:
            raise RuntimeError("StackedEncoder1DBlock runs on CUDA (sm_100a) only; there is no CPU fallback")
        eng = self.engine_for(params, tuple(x.shape), mask, bool(train), r=self.r, prop_attn=self._extra.get("prop_attn", True),
                              dropout_seed=_seed_of(dropout_rng))
        eng.forward(x.contiguous())
        self.last_size = eng.final_size()
        return eng.final_x().clone()


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def _tree_stack(trees):
    if isinstance(trees[0], dict):
        return {k: _tree_stack([t[k] for t in trees]) for k in trees[0]}
    return np.stack([_np(t) for t in trees], axis=0)


def flax_tree_to_layers(stacked: Dict[str, Any], attn_name: str, num_blocks: int):
    """Unstack axis 0 (nn.scan's `variable_axes={'params': 0}`, attention.py:103-109) into the per-layer dicts the
    executor loads: kernels flattened to [in, out]."""
    out = []
    for l in range(num_blocks):
        g = lambda *path: _np(_get(stacked, path))[l]  # noqa: E731
        a = attn_name
        c = g(a, "query", "kernel").shape[0]
        lay = dict(ln1_scale=g("LayerNorm_0", "scale"), ln1_bias=g("LayerNorm_0", "bias"),
                   ln2_scale=g("LayerNorm_1", "scale"), ln2_bias=g("LayerNorm_1", "bias"),
                   wq=g(a, "query", "kernel").reshape(c, -1), wk=g(a, "key", "kernel").reshape(c, -1),
                   wv=g(a, "value", "kernel").reshape(c, -1), bq=g(a, "query", "bias").reshape(-1),
                   bk=g(a, "key", "bias").reshape(-1), bv=g(a, "value", "bias").reshape(-1),
                   wo=g(a, "out", "kernel").reshape(-1, g(a, "out", "kernel").shape[-1]), bo=g(a, "out", "bias"),
                   w1=g("MLPBlock_0", "Dense_0", "kernel"), b1=g("MLPBlock_0", "Dense_0", "bias"),
                   w2=g("MLPBlock_0", "Dense_1", "kernel"), b2=g("MLPBlock_0", "Dense_1", "bias"))
        out.append(lay)
    return out


def _get(tree, path):
    for p in path:
        tree = tree[p]
    return tree
