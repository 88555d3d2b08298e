"""Mirror of multi_modal_transformers/action_heads: the heads that sit on the readout rows of the block stack."""
from .categorical import CategoricalActionHead, ce_loss  # noqa: F401
from .continuous import ContinuousActionHead, l2_loss  # noqa: F401
from .diffusion import DiffusionActionHead, cosine_beta_schedule  # noqa: F401
