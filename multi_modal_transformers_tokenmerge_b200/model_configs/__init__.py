"""Config loading for the block stack: the YAML schema of the reference's model_configs/attention_blocks/*.yaml
(plain PyYAML here; hydra's `compose` yields the same mapping for these files)."""
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
