"""Minimal stand-ins for the two substrates the reference's block code leans on, so that its module API survives
without Flax / Hydra (neither is installable in this image: SURVEY.md 0.4):

  * `instantiate` / `call` over `_target_` config nodes (hydra.utils, used at attention.py:32-37,58-67) for the
    handful of targets vanilla_decoder.yaml names;
  * a functional `Module` protocol with Flax's shape: `variables = m.init(rng, *args)`, `y = m.apply(variables,
    *args)`, parameters in a nested dict under "params" with Flax's names.

Host-side configuration logic only -- no arithmetic happens here.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Dict, Optional

import numpy as np


# ------------------------------------------------------------------------------------------------ config nodes
@dataclass
class LayerNormSpec:           # flax.linen.LayerNorm                      vanilla_decoder.yaml:7-13
    epsilon: float = 1e-6
    axis: int = 2              # 1 = tokens (reduction_axes: [1], what the yaml says), 2 = features


@dataclass
class DropoutSpec:             # flax.linen.Dropout                        vanilla_decoder.yaml:15-17,48-50
    rate: float = 0.0


@dataclass
class DenseSpec:               # flax.linen.Dense                          vanilla_decoder.yaml:35-42,52-59
    features: int = 0
    use_bias: bool = True
    kernel_init: str = "lecun_normal"
    bias_init: str = "zeros"


@dataclass
class AttentionSpec:           # flax.linen.SelfAttention / MultiHeadDotProductAttention / ToMe...   yaml:19-31
    num_heads: int = 1
    qkv_features: Optional[int] = None
    out_features: Optional[int] = None
    dropout_rate: float = 0.0
    broadcast_dropout: bool = True
    use_bias: bool = True
    decode: bool = False
    normalize_qk: bool = False
    kernel_init: str = "lecun_normal"
    bias_init: str = "zeros"
    tome: bool = False


_INITS = {"flax.linen.initializers.he_normal": "he_normal", "flax.linen.initializers.normal": "normal",
          "flax.linen.initializers.zeros_init": "zeros", "flax.linen.initializers.lecun_normal": "lecun_normal",
          "flax.linen.initializers.xavier_uniform": "xavier_uniform"}
_ATTN_TARGETS = {"flax.linen.SelfAttention": False, "flax.linen.MultiHeadDotProductAttention": False,
                 "flax.linen.MultiHeadAttention": False,
                 "multi_modal_transformers.attention_blocks.tome_attention.ToMeMultiHeadDotProductAttention": True}
_ACTIVATIONS = {"flax.linen.relu": "relu", "jax.nn.relu": "relu", "flax.linen.activation.relu": "relu"}


def _init_name(node, default):
    if node is None:
        return default
    if isinstance(node, str):
        return node
    t = node.get("_target_")
    if t not in _INITS:
        raise ValueError(f"unsupported initializer {t!r} (supported: {sorted(_INITS)})")
    return _INITS[t]


def instantiate(node: Dict[str, Any]):
    """hydra.utils.instantiate for the `_target_`s of vanilla_decoder.yaml -> a spec object."""
    if not isinstance(node, dict) or "_target_" not in node:
        raise ValueError(f"config node without _target_: {node!r}")
    t = node["_target_"]
    if t == "flax.linen.LayerNorm":
        red = node.get("reduction_axes", -1)
        red = list(red) if isinstance(red, (list, tuple)) else [red]
        feat = node.get("feature_axes", -1)
        feat = list(feat) if isinstance(feat, (list, tuple)) else [feat]
        if feat not in ([-1], [2]):
            raise ValueError(f"LayerNorm feature_axes {feat} unsupported (scale/bias per feature: [-1])")
        if red == [1]:
            axis = 1
        elif red in ([-1], [2]):
            axis = 2
        else:
            raise ValueError(f"LayerNorm reduction_axes {red} unsupported ([1] tokens or [-1] features)")
        return LayerNormSpec(float(node.get("epsilon", 1e-6)), axis)
    if t == "flax.linen.Dropout":
        return DropoutSpec(float(node.get("rate", 0.0)))
    if t == "flax.linen.Dense":
        return DenseSpec(int(node["features"]), bool(node.get("use_bias", True)), _init_name(node.get("kernel_init"), "lecun_normal"),
                         _init_name(node.get("bias_init"), "zeros"))
    if t in _ATTN_TARGETS:
        if node.get("decode", False):
            raise NotImplementedError("decode=True (autoregressive cache, tome_attention.py:184-236) is outside the training path")
        if node.get("normalize_qk", False):
            raise NotImplementedError("normalize_qk=True (tome_attention.py:166-180) is not implemented")
        return AttentionSpec(int(node["num_heads"]), node.get("qkv_features"), node.get("out_features"),
                             float(node.get("dropout_rate", 0.0)), bool(node.get("broadcast_dropout", True)),
                             bool(node.get("use_bias", True)), False, False, _init_name(node.get("kernel_init"), "lecun_normal"),
                             _init_name(node.get("bias_init"), "zeros"), _ATTN_TARGETS[t])
    raise ValueError(f"unsupported _target_ {t!r}")


def call(node: Dict[str, Any]) -> str:
    """hydra.utils.call on the `_partial_` activation node (attention.py:33, yaml:44-46) -> activation name."""
    t = node.get("_target_") if isinstance(node, dict) else node
    if t not in _ACTIVATIONS:
        raise NotImplementedError(f"activation {t!r}: the fused GEMM epilogue implements relu (vanilla_decoder.yaml:46)")
    return _ACTIVATIONS[t]


# ------------------------------------------------------------------------------------------------ initialisers
def make_init(name: str):
    """-> f(rng: np.random.Generator, shape, fan_in, fan_out) -> np.float32 array (flax.linen.initializers)."""
    def f(rng, shape, fan_in=None, fan_out=None):
        if name == "zeros":
            return np.zeros(shape, np.float32)
        if name == "normal":
            return (rng.standard_normal(shape) * 0.01).astype(np.float32)      # stddev 1e-2 default
        if name == "he_normal":     # variance_scaling(2.0, fan_in, truncated_normal) -- plain normal here
            return (rng.standard_normal(shape) * math.sqrt(2.0 / fan_in)).astype(np.float32)
        if name == "lecun_normal":
            return (rng.standard_normal(shape) * math.sqrt(1.0 / fan_in)).astype(np.float32)
        if name == "xavier_uniform":
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            return rng.uniform(-lim, lim, shape).astype(np.float32)
        raise ValueError(name)
    return f


def as_rng(rng) -> np.random.Generator:
    if isinstance(rng, np.random.Generator):
        return rng
    if isinstance(rng, dict):   # flax style {"params": key, "dropout": key}
        rng = rng.get("params", 0)
    return np.random.default_rng(int(np.asarray(rng).ravel()[-1]) if not isinstance(rng, int) else rng)


def merge_param(name: str, a, b):
    """flax.linen.merge_param (attention.py:54-55): exactly one of the attribute / call argument must be set."""
    if a is None and b is None:
        raise ValueError(f"Parameter {name!r} must be passed to the constructor or at call time.")
    if a is not None and b is not None:
        raise ValueError(f"Parameter {name!r} was passed to the constructor and at call time. Should be passed just once.")
    return a if b is None else b


class Module:
    """`init(rng, *args, **kw) -> {"params": tree}` and `apply(variables, *args, **kw)`; subclasses implement
    `_init(rng, *args, **kw) -> tree` and `_apply(params, *args, **kw)`."""

    def init(self, rng, *args, **kw):
        return {"params": self._init(as_rng(rng), *args, **kw)}

    def apply(self, variables, *args, rngs=None, **kw):
        if "params" not in variables:
            raise ValueError('variables must hold a "params" collection')
        if rngs is not None and "dropout" in rngs:
            kw.setdefault("dropout_rng", rngs["dropout"])
        return self._apply(variables["params"], *args, **kw)
