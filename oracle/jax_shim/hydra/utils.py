"""hydra.utils.instantiate / call for plain-dict config nodes: import `_target_`, pass the other keys as keyword
arguments (`_partial_` -> functools.partial).  Nested nodes are instantiated first when `_recursive_` is true."""
import functools
import importlib


def _resolve(path):
    parts = path.split(".")
    for i in range(len(parts) - 1, 0, -1):      # longest importable prefix, then attribute access
        try:
            obj = importlib.import_module(".".join(parts[:i]))
        except ModuleNotFoundError:
            continue
        for name in parts[i:]:
            obj = getattr(obj, name)
        return obj
    raise ImportError(path)


def instantiate(cfg, *args, _recursive_=True, **extra):
    kw = {k: v for k, v in cfg.items() if not k.startswith("_")}
    if _recursive_:
        kw = {k: instantiate(v) if isinstance(v, dict) and "_target_" in v else v for k, v in kw.items()}
    kw.update(extra)
    target = _resolve(cfg["_target_"])
    if cfg.get("_partial_"):
        return functools.partial(target, *args, **kw)
    return target(*args, **kw)


call = instantiate
