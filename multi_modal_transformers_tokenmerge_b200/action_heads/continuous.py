"""Mirror of multi_modal_transformers/action_heads/continuous.py on the sm_100a head kernel (csrc/head.cu).

    ContinuousActionHead(max_action, attention_pooling, dense)(readouts) -> [B, 1, action_dim]     continuous.py:12-25
    l2_loss(head, variables, readouts, actions) -> (per-row loss [B], mean)                        octo.py:157-165, 253-263

`dense` is the `flax.linen.Dense` config node of the head; `attention_pooling` is accepted and ignored exactly as the
reference ignores it (its use is commented out at continuous.py:18-19).  Parameters: {"Dense_0": {"kernel", "bias"}}.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from .. import _lib as L
from .. import ops
from ..attention_blocks._module import Module, instantiate, make_init


def _f32(t):
    return t if isinstance(t, torch.Tensor) else torch.as_tensor(t, dtype=torch.float32, device="cuda")


class ContinuousActionHead(Module):
    def __init__(self, max_action: float, attention_pooling: Optional[Dict[str, Any]], dense: Dict[str, Any]):
        self.max_action, self.attention_pooling, self.dense = max_action, attention_pooling, dense

    def _init(self, rng, readouts):
        d = instantiate(self.dense)
        c = readouts.shape[-1]
        p = {"kernel": make_init(d.kernel_init)(rng, (c, d.features), c, d.features)}
        if d.use_bias:
            p["bias"] = make_init(d.bias_init)(rng, (d.features,))
        return {"Dense_0": p}

    def _run(self, params, readouts, actions=None, keep_for_backward=False):
        d = instantiate(self.dense)
        w = _f32(params["Dense_0"]["kernel"]).to(readouts.device).contiguous()
        b = params["Dense_0"].get("bias")
        b = None if b is None else _f32(b).to(readouts.device).contiguous()
        if readouts.dim() != 3:
            raise ValueError("readouts must be [batch, readout, embedding] (continuous.py:17 reduces axis -2)")
        if tuple(w.shape) != (readouts.shape[-1], d.features):
            raise ValueError(f"Dense kernel {tuple(w.shape)} does not match ({readouts.shape[-1]}, {d.features})")
        return ops.action_head_fwd(readouts.contiguous(), w, b, kind=L.HEAD_CONTINUOUS_L2, max_action=float(self.max_action),
                                   actions=actions, keep_for_backward=keep_for_backward)

    def _apply(self, params, readouts, dropout_rng=None):
        out, _, _ = self._run(params, readouts)
        return out                                                     # [B, 1, action_dim] (continuous.py:22)


def l2_loss(head: ContinuousActionHead, variables, readouts, actions):
    """Octo.compute_l2_loss on already-computed readouts: (sum_a (pred - action)^2 per batch row [B], its mean)."""
    _, loss, _ = head._run(variables["params"], readouts, actions=actions.contiguous())
    return loss[1:], loss[0]
