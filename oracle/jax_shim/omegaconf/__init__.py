"""omegaconf stand-in: the reference only uses DictConfig as a type annotation (test infrastructure, see ../README.md)."""
DictConfig = dict


class OmegaConf:   # imported by tokenizers/images/image_tokenizer.py, never used by the code paths the generator executes
    pass
