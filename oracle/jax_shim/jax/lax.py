import numpy as _np
from .numpy import _wrap


def top_k(x, k):
    idx = _np.argsort(-_np.asarray(x), axis=-1, kind="stable")[..., :k]
    return _wrap(_np.take_along_axis(_np.asarray(x), idx, axis=-1)), _wrap(idx)


def dynamic_slice_in_dim(x, start, size, axis=0):
    sl = [slice(None)] * _np.ndim(x)
    sl[axis] = slice(start, start + size)
    return _wrap(_np.asarray(x)[tuple(sl)])
