"""jax.numpy stand-in on numpy (fp32 default, functional .at[] updates)."""
import numpy as _np

inf = _np.inf
pi = _np.pi
float32 = _np.float32
int32 = _np.int32
bool_ = _np.bool_


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v):
        out = _np.array(self.arr, copy=True)
        out[self.idx] = v
        return _wrap(out)

    def add(self, v):
        out = _np.array(self.arr, copy=True)
        _np.add.at(out, self.idx, v)  # accumulates over repeated indices, like XLA scatter-add
        return _wrap(out)


class ndarray(_np.ndarray):
    @property
    def at(self):
        return _At(self)


array_t = ndarray


def _wrap(a):
    a = _np.asarray(a)
    if a.dtype == _np.float64:
        a = a.astype(_np.float32)  # jax default: x64 disabled
    if a.dtype == _np.int64:
        a = a.astype(_np.int32)
    return a.view(ndarray)


def asarray(a, dtype=None):
    return _wrap(_np.array(a, dtype=dtype, copy=True))


array = asarray


def _lift(name):
    f = getattr(_np, name)

    def g(*a, **k):
        return _wrap(f(*a, **k))

    g.__name__ = name
    return g


for _n in ("matmul", "swapaxes", "take_along_axis", "concatenate", "ones_like", "zeros_like", "arange", "ones",
           "zeros", "hstack", "vstack", "squeeze", "take", "ravel", "sum", "expand_dims", "repeat", "tile",
           "where", "stack", "sqrt", "log", "exp", "maximum", "minimum", "abs", "mean", "tanh", "digitize", "cos", "sin", "clip", "prod",
           "floor", "reshape"):
    globals()[_n] = _lift(_n)


def argsort(a, axis=-1, kind=None, order=None, stable=True, descending=False):
    assert not descending
    return _wrap(_np.argsort(a, axis=axis, kind="stable"))


def argmax(a, axis=None):
    return _wrap(_np.argmax(a, axis=axis))


class linalg:
    @staticmethod
    def norm(x, axis=None, keepdims=False):
        x = _np.asarray(x)
        return _wrap(_np.sqrt(_np.sum(x * x, axis=axis, keepdims=keepdims)))


def linspace(start, stop, num=50, dtype=None):
    """jnp.linspace with x64 disabled: computed in fp32 (start + i * step, last point = stop)."""
    return _wrap(_np.linspace(_np.float32(start), _np.float32(stop), num, dtype=_np.float32))
