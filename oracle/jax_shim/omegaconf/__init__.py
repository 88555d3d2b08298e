"""omegaconf stand-in: the reference only uses DictConfig as a type annotation (test infrastructure, see ../README.md)."""
DictConfig = dict
