"""numpy-backed stand-in for the few jax symbols the reference's token_compression.py /
token_sequencer.py touch. Test infrastructure only (see ../README.md)."""
import functools
from . import numpy  # noqa: F401
from . import lax, typing, random, debug  # noqa: F401

Array = numpy.ndarray


def jit(fn=None, static_argnums=None, static_argnames=None, **kw):
    if fn is None:
        return functools.partial(jit, static_argnums=static_argnums, static_argnames=static_argnames)
    return fn


def vmap(fn, in_axes=0, out_axes=0):
    import numpy as _np

    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(a.shape[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = [fn(*[a if ax is None else _np.take(a, i, axis=ax) for a, ax in zip(args, axes)]) for i in range(n)]
        if isinstance(outs[0], tuple):   # functions returning several arrays: map each
            return tuple(numpy._wrap(_np.stack([o[j] for o in outs], axis=out_axes)) for j in range(len(outs[0])))
        return numpy._wrap(_np.stack(outs, axis=out_axes))

    return mapped
