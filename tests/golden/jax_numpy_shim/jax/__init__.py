"""NumPy-backed stand-in for the slice of the JAX API that the reference's EK1 path uses.

jax/jaxlib (<= 0.3.1, which the reference needs) cannot be installed in this image.  This shim lets the reference's OWN
source files (src/pnmol/{white,latent,pdefilter}.py, base/{sqrt,iwp,stacked_ssm,rv}.py, odetools/step.py) execute
unmodified on float64 NumPy/SciPy -- the same LAPACK routines (dgeqrf, dtrtrs, dpotrf) that jaxlib's CPU backend
dispatches to -- so that golden vectors come from the reference's code rather than from a restatement of it.
Only used by tests/golden/make_reference_golden.py; nothing in the product or in the tests imports it at run time.
"""
import functools

import numpy as _np

from . import numpy, scipy  # noqa: F401
from . import config as _config_module
from .config import config  # noqa: F401


class Array:  # scipy's array-API helpers ask sys.modules["jax"] for this type; nothing here is an instance of it
    pass


def jit(fun=None, *, static_argnums=None, static_argnames=None, **_):
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(_np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = [fun(*[_np.take(a, i, axis=ax) if ax is not None else a for a, ax in zip(args, axes)]) for i in range(n)]
        if isinstance(outs[0], tuple):
            return tuple(_np.stack([o[k] for o in outs], axis=out_axes) for k in range(len(outs[0])))
        return _np.stack(outs, axis=out_axes)

    return mapped


def _no_autodiff(*_a, **_k):
    def fail(*_a2, **_k2):
        raise NotImplementedError("the NumPy shim has no autodiff: supply f/df and the discretised operators explicitly")

    return fail


grad = jacfwd = jacrev = jacobian = hessian = value_and_grad = _no_autodiff


class _Ops:
    @staticmethod
    def index_update(x, idx, y):
        x = _np.array(x, copy=True)
        x[idx] = y
        return x

    class _Index:
        def __getitem__(self, item):
            return item

    index = _Index()


ops = _Ops()
partial = functools.partial
