"""Config loading for the block stack and the action heads: the YAML schema of the reference's
model_configs/attention_blocks/*.yaml and model_configs/action_heads/*.yaml (plain PyYAML here; hydra's `compose` yields the
same mapping for these files)."""
from __future__ import annotations

import os
from typing import Any, Dict

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name_or_path: str) -> Dict[str, Any]:
    """`load("attention_blocks/tome_decoder_octo_small")` or a path to any YAML with the vanilla_decoder.yaml schema."""
    p = name_or_path
    if not os.path.exists(p):
        p = os.path.join(HERE, name_or_path + ("" if name_or_path.endswith(".yaml") else ".yaml"))
    with open(p) as f:
        return yaml.safe_load(f)


def build_stack(cfg: Dict[str, Any]):
    """What octo.py:80 does with `instantiate(config.attention_blocks...)`: -> StackedEncoder1DBlock for this config.
    A ToMe block target (or a non-zero `tome_r`) selects the ToMe stack; the vanilla target gives the plain stack."""
    from ..attention_blocks import attention, tome_attention

    target = cfg["encoder_1d_block"].get("_target_", "")
    r = int(cfg.get("tome_r", 0))
    if target.endswith("ToMeEncoder1DBlock") or r > 0:
        return tome_attention.StackedEncoder1DBlock(cfg["num_blocks"], cfg["encoder_1d_block"], tome_r=r,
                                                    prop_attn=bool(cfg.get("prop_attn", True)))
    return attention.StackedEncoder1DBlock(cfg["num_blocks"], cfg["encoder_1d_block"])


_HEAD_TARGETS = {"ContinuousActionHead": ("continuous", ("max_action", "attention_pooling", "dense")),
                 "CategoricalActionHead": ("categorical", ("num_bins", "max_action", "action_space_dim", "dense")),
                 "DiffusionActionHead": ("diffusion", ("diffusion_steps", "attention_pooling", "denoising_model", "rng_collection"))}


def build_action_head(node: Dict[str, Any]):
    """What octo.py:82-86 does with `instantiate(action_head...)`: the head module named by the node's `_target_`
    (multi_modal_transformers.action_heads.{continuous,categorical,diffusion}.*ActionHead), built from the node's own keys.
    Accepts the head node itself or the one-entry mapping a head YAML holds (`diffusion_action_head: {...}`)."""
    from .. import action_heads

    if "_target_" not in node and len(node) == 1:
        node = next(iter(node.values()))
    cls = str(node.get("_target_", "")).rsplit(".", 1)[-1]
    if cls not in _HEAD_TARGETS:
        raise ValueError(f"unsupported action head _target_ {node.get('_target_')!r} (supported: {sorted(_HEAD_TARGETS)})")
    _, keys = _HEAD_TARGETS[cls]
    kw = {k: node[k] for k in keys if k in node}
    if "attention_pooling" in keys:
        kw.setdefault("attention_pooling", None)
    return getattr(action_heads, cls)(**kw)


def build_image_tokenizer(node: Dict[str, Any], **overrides):
    """What the reference does with `instantiate(config.tokenizers.images.encoder)`: the ImageTokenizer named by the node
    (the `encoder` mapping of gato_resnet.yaml, or the YAML's top level holding it).  Hydra interpolations of the reference file
    (`${tokenizers.images.encoder.position_interval}`, `${dtype}`) are resolved the way its composed config resolves them:
    num_embeddings = position_interval; dtype keys are dropped (activations are bf16, parameters fp32 here)."""
    from ..tokenizers.images import ImageTokenizer

    if "_target_" not in node and "encoder" in node:
        node = node["encoder"]
    if not str(node.get("_target_", "")).endswith("ImageTokenizer"):
        raise ValueError(f"unsupported image tokenizer _target_ {node.get('_target_')!r}")

    def resolve(v):
        if isinstance(v, dict):
            return {k: resolve(x) for k, x in v.items() if k not in ("dtype", "param_dtype")}
        if isinstance(v, str) and v.startswith("${") and v.endswith("position_interval}"):
            return int(node["position_interval"])
        return v

    kw = {k: resolve(v) for k, v in node.items() if k != "_target_"}
    kw.update(overrides)
    return ImageTokenizer(**kw)
