"""Mirror of multi_modal_transformers/action_heads/categorical.py on the sm_100a head kernel (csrc/head.cu).

    assign_bins(input_data, bounds, num_bins, bin_strategy="uniform")                              categorical.py:12-22
    CategoricalActionHead(num_bins, max_action, action_space_dim, dense)(readouts) -> logits [B, action, num_bins]   :24-40
    ce_loss(head, variables, readouts, actions) -> (per-row loss summed over actions [B], mean over (B, action))
                                                                                                   octo.py:178-190, 292-303

The readouts are laid out "(action timestep)": readout i feeds action i // timesteps.  The label arithmetic is the
reference's, literally: `jnp.digitize` is 1-based, so a value in bin k is labelled k + 1 and the top bin (and anything
above the range) gets an all-zero one-hot row, whose cross-entropy is 0.
"""
from __future__ import annotations

from typing import Any, Dict

import numpy as np
import torch

from .. import _lib as L
from .. import ops
from ..attention_blocks._module import Module, instantiate, make_init
from .continuous import _f32


def assign_bins(input_data, bounds, num_bins, bin_strategy="uniform"):
    """Host-side helper (index arithmetic; the kernel recomputes the same bins on the device for the loss)."""
    if bin_strategy != "uniform":
        raise NotImplementedError
    bins = np.linspace(np.float32(bounds[0]), np.float32(bounds[1]), num_bins + 1, dtype=np.float32)
    x = input_data.detach().cpu().numpy() if isinstance(input_data, torch.Tensor) else np.asarray(input_data)
    return np.digitize(x.astype(np.float32), bins)


class CategoricalActionHead(Module):
    def __init__(self, num_bins: int, max_action: float, action_space_dim: int, dense: Dict[str, Any]):
        self.num_bins, self.max_action, self.action_space_dim, self.dense = num_bins, max_action, action_space_dim, dense

    def _init(self, rng, readouts):
        d = instantiate(self.dense)
        c = readouts.shape[-1]
        p = {"kernel": make_init(d.kernel_init)(rng, (c, d.features), c, d.features)}
        if d.use_bias:
            p["bias"] = make_init(d.bias_init)(rng, (d.features,))
        return {"Dense_0": p}

    def _run(self, params, readouts, actions=None, keep_for_backward=False):
        d = instantiate(self.dense)
        if d.features != self.num_bins:
            raise ValueError(f"dense.features ({d.features}) must equal num_bins ({self.num_bins}): the logits are compared "
                             "with one_hot(bins, num_bins) (octo.py:183-187)")
        if readouts.dim() != 3 or readouts.shape[1] % self.action_space_dim:
            raise ValueError(f"readouts {tuple(readouts.shape)} do not split as '(action timestep)' with action = "
                             f"{self.action_space_dim} (categorical.py:32-36)")
        w = _f32(params["Dense_0"]["kernel"]).to(readouts.device).contiguous()
        b = params["Dense_0"].get("bias")
        b = None if b is None else _f32(b).to(readouts.device).contiguous()
        return ops.action_head_fwd(readouts.contiguous(), w, b, kind=L.HEAD_CATEGORICAL_CE, max_action=float(self.max_action),
                                   groups=self.action_space_dim, actions=actions, keep_for_backward=keep_for_backward)

    def _apply(self, params, readouts, dropout_rng=None):
        out, _, _ = self._run(params, readouts)
        return out                                                     # logits [B, action, num_bins]


def ce_loss(head: CategoricalActionHead, variables, readouts, actions):
    """Octo.compute_ce_loss on already-computed readouts: (cross-entropy summed over the action axis [B], mean over
    (B, action) as categorical_train_step takes it)."""
    _, loss, _ = head._run(variables["params"], readouts, actions=actions.contiguous())
    return loss[1:], loss[0]
