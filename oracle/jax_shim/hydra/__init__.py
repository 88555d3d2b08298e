"""hydra stand-in (test infrastructure, see ../README.md)."""
from . import utils  # noqa: F401
