"""Covariance kernels (host-side set-up, NumPy float64).

API of src/pnmol/kernels.py: ``k(X, Y)`` with X (N, dim) and Y (dim, K) returns the Gram
matrix (N, K); equal shapes return the diagonal; ``k1 + k2`` sums kernels;
``duplicate(k, num)`` makes a block-diagonal Gram matrix.  Derivatives needed by the
probabilistic finite differences are closed-form in one dimension (``derivative``)
instead of JAX autodiff.
"""
import numpy as np
import scipy.linalg


class Kernel:
    def pairwise(self, x, y):
        raise NotImplementedError

    def __call__(self, X, Y):
        X, Y = np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64)
        if X.ndim == Y.ndim <= 1:  # kernels.py:23-24
            return self.pairwise(X, Y)
        if X.shape == Y.shape:  # diagonal, kernels.py:28-29
            return self._eval(X, Y)
        return self._eval(X[:, None, :], Y.T[None, :, :])  # full Gram, kernels.py:31-32

    def _eval(self, X, Y):
        raise NotImplementedError

    def __add__(self, other):
        return _Sum(self, other)

    def __str__(self):
        return f"{type(self).__name__}()"

    def derivative(self, which, x, y):
        """1-D derivative of k at scalar arrays (x, y): which in {dx, dxx, dxy, dxxyy}."""
        raise NotImplementedError(f"{type(self).__name__} has no closed-form derivative {which}")


class _Sum(Kernel):
    def __init__(self, a, b):
        self.a, self.b = a, b

    def pairwise(self, x, y):
        return self.a.pairwise(x, y) + self.b.pairwise(x, y)

    def _eval(self, X, Y):
        return self.a._eval(X, Y) + self.b._eval(X, Y)

    def derivative(self, which, x, y):
        return self.a.derivative(which, x, y) + self.b.derivative(which, x, y)


class _Radial(Kernel):
    def __init__(self, *, output_scale=1.0, input_scale=1.0):
        self.output_scale, self.input_scale = output_scale, input_scale

    @staticmethod
    def _sqdist(X, Y):
        diff = X - Y
        return np.sum(diff * diff, axis=-1)

    def pairwise(self, x, y):
        return self._eval(np.atleast_1d(x), np.atleast_1d(y))


class SquareExponential(_Radial):
    """kernels.py:107-111."""

    def _eval(self, X, Y):
        return self.output_scale ** 2 * np.exp(-self._sqdist(X, Y) * self.input_scale ** 2 / 2.0)

    def derivative(self, which, x, y):
        q, u = self.input_scale ** 2, x - y
        k = self.output_scale ** 2 * np.exp(-q * u * u / 2.0)
        poly = {"dx": -q * u, "dxx": q * q * u * u - q, "dxy": q - q * q * u * u,
                "dxxyy": 3 * q ** 2 - 6 * q ** 3 * u ** 2 + q ** 4 * u ** 4}[which]
        return poly * k


class Matern52(_Radial):
    """kernels.py:114-124.  Autodiff of the reference yields NaN at x == y; derivatives here
    return NaN there too and the discretisation substitutes the reference's constants
    (discretize.py:184-197)."""

    def _eval(self, X, Y):
        a = np.sqrt(5.0 * self._sqdist(X, Y) * self.input_scale ** 2)
        return self.output_scale ** 2 * (1 + a + a ** 2.0 / 3.0) * np.exp(-a)

    def derivative(self, which, x, y):
        u = np.asarray(x - y, dtype=np.float64)
        c = np.sqrt(5.0) * self.input_scale
        a = c * np.abs(u)
        s2, ex = self.output_scale ** 2, np.exp(-a)
        val = {"dx": -s2 / 3.0 * c * np.sign(u) * a * (1 + a) * ex,
               "dxx": -s2 / 3.0 * c ** 2 * (1 + a - a * a) * ex,
               "dxy": s2 / 3.0 * c ** 2 * (1 + a - a * a) * ex,
               "dxxyy": -s2 / 3.0 * c ** 4 * (-a * a + 5 * a - 3) * ex}[which]
        return np.where(u == 0.0, np.nan, val)


class Polynomial(Kernel):
    """kernels.py:127-144: (x^T y + const)^order."""

    def __init__(self, *, order=2, const=1.0):
        self.order, self.const = order, const

    def pairwise(self, x, y):
        return (np.dot(x, y) + self.const) ** self.order

    def _eval(self, X, Y):
        return (np.sum(X * Y, axis=-1) + self.const) ** self.order

    def derivative(self, which, x, y):
        p, b = self.order, x * y + self.const

        def pw(e):
            return b ** e if e >= 0 else np.zeros_like(b)

        if which == "dx":
            return p * y * pw(p - 1)
        if which == "dxx":
            return p * (p - 1) * y ** 2 * pw(p - 2)
        if which == "dxy":
            return p * pw(p - 1) + p * (p - 1) * x * y * pw(p - 2)
        if which == "dxxyy":
            return p * (p - 1) * (2.0 * pw(p - 2) + 4.0 * x * y * (p - 2) * pw(p - 3)
                                  + x ** 2 * y ** 2 * (p - 2) * (p - 3) * pw(p - 4))
        raise KeyError(which)


class WhiteNoise(Kernel):
    """kernels.py:147-157."""

    def __init__(self, *, output_scale=1.0):
        self.output_scale = output_scale

    def pairwise(self, x, y):
        return self.output_scale ** 2 * float(np.all(np.asarray(x) == np.asarray(y)))

    def _eval(self, X, Y):
        return self.output_scale ** 2 * np.all(X == Y, axis=-1).astype(np.float64)


class _StackedKernel(Kernel):
    """kernels.py:160-175."""

    def __init__(self, *, kernel_list):
        self.kernel_list = kernel_list

    def __call__(self, X, Y):
        grams = [k(X, Y) for k in self.kernel_list]
        if np.shape(X) == np.shape(Y):
            return np.concatenate(grams)
        return scipy.linalg.block_diag(*grams)


def duplicate(kernel, num):
    """kernels.py:178-183."""
    return _StackedKernel(kernel_list=[kernel] * num)


# ------------------------------------------------------------------------- batched, on the device
def _device_kernel_table(kernel_list):
    """(kind, params [B, 4], white-noise variance) of a list of kernels of ONE type (optionally `+ WhiteNoise`)."""
    kinds, rows, white = set(), [], set()
    for k in kernel_list:
        w = 0.0
        if isinstance(k, _Sum):
            base, wn = (k.a, k.b) if isinstance(k.b, WhiteNoise) else (k.b, k.a)
            if not isinstance(wn, WhiteNoise):
                raise NotImplementedError("device kernels: only `kernel + WhiteNoise` sums")
            k, w = base, wn.output_scale ** 2
        if isinstance(k, SquareExponential):
            kinds.add(0); rows.append((k.input_scale, k.output_scale, 0.0, 0.0))
        elif isinstance(k, Matern52):
            kinds.add(1); rows.append((k.input_scale, k.output_scale, 0.0, 0.0))
        elif isinstance(k, Polynomial):
            kinds.add(2); rows.append((1.0, 1.0, float(k.order), float(k.const)))
        else:
            raise NotImplementedError(f"device kernels: {type(k).__name__} has no closed-form device implementation")
        white.add(w)
    if len(kinds) != 1 or len(white) != 1:
        raise ValueError("a batch must hold kernels of one type with one white-noise scale")
    return kinds.pop(), np.asarray(rows, dtype=np.float64), white.pop()


def gram_cholesky_batched(kernel_list, mesh_points, *, data=None, nugget=0.0, device=None):
    """``cholesky(k(X, X.T))`` for every kernel of the list in one launch (white.py:82-94 initialize_iwp, batched over
    kernel hyper-parameters), and -- with ``data`` -- the log-likelihood of kernels.py:203-211 under each Gram matrix.

    Returns torch CUDA tensors ``chol [B, d, d]`` (lower), ``status [B]`` (1 = not positive definite) and
    ``loglik [B]`` or None."""
    import torch

    from . import _lib

    kind, params, white = _device_kernel_table(kernel_list)
    X = np.asarray(mesh_points, dtype=np.float64)
    if X.ndim == 2:
        if X.shape[1] != 1:
            raise NotImplementedError("device kernels are 1-D")
        X = X[:, 0]
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    dev = torch.device("cuda", idx if idx is not None else torch.cuda.current_device())
    lib = _lib.load()
    B, d = len(params), len(X)
    par = torch.as_tensor(params, device=dev)
    pts = torch.as_tensor(X, device=dev)
    y = None if data is None else torch.as_tensor(np.asarray(data, dtype=np.float64).reshape(-1), device=dev)
    chol = torch.empty((B, d, d), dtype=torch.float64, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    ll = None if data is None else torch.empty(B, dtype=torch.float64, device=dev)
    _lib.check(lib.pnmol_b200_gram_cholesky(kind, _lib.ptr(par), B, _lib.ptr(pts), d, float(white + nugget), _lib.ptr(y),
                                            _lib.ptr(chol), _lib.ptr(ll), _lib.ptr(status), dev.index,
                                            _lib.current_stream(dev)))
    return chol, status, ll


def mle_input_scale(*, mesh_points, data, kernel_type, input_scale_trials, nugget=0.0):
    """kernels.py:186-200: the trial input scale with the largest Gaussian log-likelihood of ``data``; all trials are
    factored in one launch (Cholesky instead of the reference's LU-based solve/det: same value for a positive
    definite Gram matrix)."""
    trials = np.asarray(input_scale_trials, dtype=np.float64)
    _, _, ll = gram_cholesky_batched([kernel_type(input_scale=float(s)) for s in trials], mesh_points, data=data, nugget=nugget)
    return trials[int(np.argmax(np.nan_to_num(ll.cpu().numpy(), nan=-np.inf)))]
