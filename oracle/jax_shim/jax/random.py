import numpy as _np
from .numpy import _wrap
def PRNGKey(s): return _np.random.default_rng(s)
def normal(key, shape): return _wrap(key.standard_normal(shape).astype(_np.float32))
