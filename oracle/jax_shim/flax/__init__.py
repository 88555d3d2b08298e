"""flax stand-in: only what token_sequencer.py imports (test infrastructure, see ../README.md)."""
from . import linen, struct  # noqa: F401
