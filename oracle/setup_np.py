"""Oracle (test infrastructure): problem set-up of the reference in NumPy float64.

Restates, for one spatial dimension, ``src/pnmol/mesh.py:86-98,132-175`` (mesh),
``src/pnmol/kernels.py:17-57,107-183`` (covariance kernels), ``src/pnmol/discretize.py:12-201``
(probabilistic finite differences), ``src/pnmol/pde/mixins.py:19-122`` (discretisation
mix-ins) and ``src/pnmol/pde/examples.py:13-357`` (problem recipes).  The reference
differentiates its kernels with JAX autodiff; here the 1-D derivatives are written in
closed form (``tests/test_oracle_setup.py`` checks them against ``torch.autograd`` in
float64).  Everything is written point-by-point (Python loops, ``numpy.linalg.solve`` per
stencil) -- slow but literal.
"""
from types import SimpleNamespace

import numpy as np
import scipy.linalg
import scipy.spatial


# ------------------------------------------------------------------ kernels (1-D)
class SE:
    """kernels.py:107-111: s^2 exp(-r^2 (x-y)^2 / 2)."""

    def __init__(self, input_scale=1.0, output_scale=1.0):
        self.r, self.s = input_scale, output_scale

    def k(self, x, y):
        return self.s ** 2 * np.exp(-self.r ** 2 * (x - y) ** 2 / 2.0)

    def dxx(self, x, y):  # d^2/dx^2
        q, u = self.r ** 2, x - y
        return (q * q * u * u - q) * self.k(x, y)

    def dxxyy(self, x, y):  # d^4/dx^2 dy^2
        q, u = self.r ** 2, x - y
        return (3 * q ** 2 - 6 * q ** 3 * u ** 2 + q ** 4 * u ** 4) * self.k(x, y)

    def dx(self, x, y):
        return -self.r ** 2 * (x - y) * self.k(x, y)

    def dxy(self, x, y):
        q, u = self.r ** 2, x - y
        return (q - q * q * u * u) * self.k(x, y)


class Matern52:
    """kernels.py:114-124.  Not differentiable at x == y: JAX autodiff returns NaN there
    and the reference substitutes Taylor constants (discretize.py:184-197) -- for ANY
    differential operator, which this class reproduces via ``nan`` results at u == 0."""

    def __init__(self, input_scale=1.0, output_scale=1.0):
        self.r, self.s = input_scale, output_scale

    def _a(self, x, y):
        return np.sqrt(5.0) * self.r * np.abs(x - y)

    def k(self, x, y):
        a = self._a(x, y)
        return self.s ** 2 * (1 + a + a ** 2 / 3.0) * np.exp(-a)

    def _guard(self, val, x, y):
        return np.where(np.asarray(x - y) == 0.0, np.nan, val)

    def dxx(self, x, y):
        a = self._a(x, y)
        c2 = 5.0 * self.r ** 2
        return self._guard(-self.s ** 2 / 3.0 * c2 * (1 + a - a * a) * np.exp(-a), x, y)

    def dxxyy(self, x, y):
        a = self._a(x, y)
        c4 = 25.0 * self.r ** 4
        return self._guard(-self.s ** 2 / 3.0 * c4 * (-a * a + 5 * a - 3) * np.exp(-a), x, y)

    def dx(self, x, y):
        a = self._a(x, y)
        c = np.sqrt(5.0) * self.r
        return self._guard(-self.s ** 2 / 3.0 * c * np.sign(x - y) * a * (1 + a) * np.exp(-a), x, y)

    def dxy(self, x, y):
        a = self._a(x, y)
        c2 = 5.0 * self.r ** 2
        return self._guard(self.s ** 2 / 3.0 * c2 * (1 + a - a * a) * np.exp(-a), x, y)

    @property
    def nan_first(self):  # discretize.py:187
        return self.r ** 2 * self.s ** 2 * 2.5 / (1.0 - 2.5)

    @property
    def nan_second(self):  # discretize.py:195-196
        return self.s ** 2 * self.r ** 4 * 3 * 2.5 ** 2 / (2.0 - 3 * 2.5 + 2.5 ** 2)


class Poly:
    """kernels.py:127-144: (x y + c)^p."""

    def __init__(self, order=2, const=1.0):
        self.p, self.c = order, const

    def k(self, x, y):
        return (x * y + self.c) ** self.p

    def dxx(self, x, y):
        p = self.p
        return p * (p - 1) * y ** 2 * (x * y + self.c) ** (p - 2) if p >= 2 else 0.0 * x * y

    def dxxyy(self, x, y):
        p, b = self.p, x * y + self.c
        out = 2.0 * _pw(b, p - 2)
        out = out + 4.0 * y * (p - 2) * x * _pw(b, p - 3)
        out = out + y ** 2 * (p - 2) * (p - 3) * x ** 2 * _pw(b, p - 4)
        return p * (p - 1) * out

    def dx(self, x, y):
        return self.p * y * _pw(x * y + self.c, self.p - 1)

    def dxy(self, x, y):
        p, b = self.p, x * y + self.c
        return p * _pw(b, p - 1) + p * (p - 1) * x * y * _pw(b, p - 2)


def _pw(b, e):
    return b ** e if e >= 0 else 0.0 * b


class White:
    """kernels.py:147-157."""

    def __init__(self, output_scale=1.0):
        self.s = output_scale

    def k(self, x, y):
        return self.s ** 2 * (np.asarray(x) == np.asarray(y)).astype(np.float64)


class Sum:
    """kernels.py:52-57 (``__add__``)."""

    def __init__(self, *parts):
        self.parts = parts

    def k(self, x, y):
        return sum(p.k(x, y) for p in self.parts)


def gram(kernel, pts, copies=1):
    """k(X, X^T) (kernels.py:17-32); ``copies`` > 1 = ``duplicate`` (kernels.py:160-183)."""
    x = np.asarray(pts, dtype=np.float64).reshape(-1)
    G = kernel.k(x[:, None], x[None, :])
    return scipy.linalg.block_diag(*([G] * copies)) if copies > 1 else G


# --------------------------------------------------------------------------- mesh
def mesh_1d(bbox, step=None, num=None):
    """mesh.py:86-98, including the floating-point floor of quirk Q6."""
    lo, hi = float(bbox[0]), float(bbox[1])
    if (step is None) == (num is None):
        raise ValueError("Provide exactly one of step or num.")
    if step is not None:
        num = int((hi - lo) / step) + 1
    return np.linspace(lo, hi, num, endpoint=True).reshape(-1, 1)


def _split_boundary(points):
    """mesh.py:141-168 with bbox read off the points (mesh.py:178-184)."""
    x = points[:, 0]
    on = np.logical_or(x == x.min(), x == x.max())
    return np.nonzero(on)[0], np.nonzero(~on)[0]


# --------------------------------------------------------------------- discretise
def stencil_weights(kernel, x, nbrs, mode, nugget=0.0):
    """discretize.py:177-201 for ``mode`` 'laplace' (dxx / dxxyy) or 'gradient' (dx / dxy)."""
    d1 = kernel.dxx if mode == "laplace" else kernel.dx
    d2 = kernel.dxxyy if mode == "laplace" else kernel.dxy
    X = np.asarray(nbrs, dtype=np.float64)
    G = kernel.k(X[:, None], X[None, :]) + nugget * np.eye(X.shape[0])
    rhs = np.array([d1(x, xj) for xj in X], dtype=np.float64)
    if isinstance(kernel, Matern52):
        rhs = np.where(np.isnan(rhs), kernel.nan_first, rhs)
    w = np.linalg.solve(G, rhs)
    top = float(d2(x, x))
    if isinstance(kernel, Matern52) and np.isnan(top):
        top = kernel.nan_second
    return w, top - w @ rhs


def fd_laplace(points, kernel, n_int=3, n_bnd=3, nugget=0.0):
    """discretize.py:12-113: dense L and diagonal E_sqrtm (holding the posterior
    *variance*, quirk Q3)."""
    N = points.shape[0]
    tree = scipy.spatial.KDTree(points)
    bnd, inner = _split_boundary(points)
    L = np.zeros((N, N))
    E = np.zeros((N, N))
    for rows, size in ((bnd, n_bnd), (inner, n_int)):
        for i in rows:
            _, idx = tree.query(points[i], k=size)
            idx = np.atleast_1d(idx)
            w, unc = stencil_weights(kernel, points[i, 0], points[idx, 0], "laplace", nugget)
            L[i, idx] = w
            E[i, i] = unc
    return L, E


def fd_neumann(points, kernel, nugget=0.0):
    """discretize.py:116-174."""
    N = points.shape[0]
    x = points[:, 0]
    wl, ul = stencil_weights(kernel, x[0], x[[0, 1]], "gradient", nugget)
    wr, ur = stencil_weights(kernel, x[-1], x[[-1, -2]], "gradient", nugget)
    Bsel = np.eye(N)[[0, 1, N - 1, N - 2]]
    diffmat = scipy.linalg.block_diag(-wl[None, :], wr[None, :])
    return diffmat @ Bsel, np.diag(np.array([ul, ur]))


def discretise(points, kernel, scales, bcond, n_int=3, n_bnd=3, nugget=0.0):
    """mixins.py:19-59 (scalar) / 66-122 (systems: one Laplacian per component, Neumann
    only -- quirk Q12)."""
    scales = np.atleast_1d(np.asarray(scales, dtype=np.float64))
    L1, E1 = fd_laplace(points, kernel, n_int, n_bnd, nugget)
    L = scipy.linalg.block_diag(*[s * L1 for s in scales])
    E = scipy.linalg.block_diag(*[s * E1 for s in scales])
    if bcond == "neumann":
        B1, R1 = fd_neumann(points, kernel, nugget)
    elif bcond == "dirichlet":
        if scales.shape[0] > 1:
            raise RuntimeError("system + Dirichlet is broken in the reference (mixins.py:112-114)")
        bnd, _ = _split_boundary(points)
        B1 = np.eye(points.shape[0])[bnd]
        R1 = np.zeros((B1.shape[0], B1.shape[0]))
    else:
        raise ValueError(bcond)
    c = scales.shape[0]
    return L, E, scipy.linalg.block_diag(*([B1] * c)), scipy.linalg.block_diag(*([R1] * c))


# ------------------------------------------------------------------------ recipes
def bell_centered(x, bbox, width=1.0):  # examples.py:347-349
    mid = 0.5 * (bbox[1] + bbox[0])
    return np.exp(-((x - mid) ** 2) / width ** 2)


def sin_bell(x):  # examples.py:356-357
    return 0.1 * np.sin(np.pi * x)


def _problem(points, kernel, scales, bcond, y0, t0, tmax, f=None, df=None, **disc):
    L, E, B, R = discretise(points, kernel, scales, bcond, **disc)
    return SimpleNamespace(L=L, E_sqrtm=E, B=B, R_sqrtm=R, y0=y0, t0=t0, tmax=tmax, f=f, df=df,
                           points=points)


def heat_1d(*, bbox=(0.0, 1.0), dx=None, num=None, t0=0.0, tmax=5.0, diffusion_rate=0.05, kernel=None,
            bcond="dirichlet", y0=None, n_int=3, n_bnd=3, nugget=0.0):
    """examples.py:13-81."""
    pts = mesh_1d(bbox, step=dx, num=num)
    kernel = kernel or SE()
    if y0 is None:
        y0 = bell_centered(pts[:, 0], bbox) * sin_bell(pts[:, 0])
    return _problem(pts, kernel, diffusion_rate, bcond, y0, t0, tmax, n_int=n_int, n_bnd=n_bnd, nugget=nugget)


def spruce_budworm_1d(*, bbox=(0.0, 1.0), dx=None, num=None, t0=0.0, tmax=10.0, diffusion_rate=1.0,
                      growth_rate=1.0, kernel=None, bcond="dirichlet", y0=None, n_int=3, n_bnd=3, nugget=0.0):
    """examples.py:251-341: f(x) = c x (1 - x), Jacobian diag(c (1 - 2x))."""
    pts = mesh_1d(bbox, step=dx, num=num)
    kernel = kernel or SE()
    if y0 is None:
        y0 = sin_bell(pts[:, 0])
    c = growth_rate
    return _problem(pts, kernel, diffusion_rate, bcond, y0, t0, tmax,
                    f=lambda t, x: c * x * (1.0 - x), df=lambda t, x: np.diag(c * (1.0 - 2.0 * x)),
                    n_int=n_int, n_bnd=n_bnd, nugget=nugget)


def sir_1d(*, bbox=(0.0, 1.0), dx=None, num=None, t0=0.0, tmax=50.0, beta=0.3, gamma=0.07, N=1000.0,
           diffusion_rates=(0.1, 0.1, 0.1), kernel=None, n_int=3, n_bnd=3, nugget=0.0):
    """examples.py:84-178 (Neumann system of three components)."""
    pts = mesh_1d(bbox, step=dx, num=num)
    kernel = kernel or SE()
    x = pts[:, 0]
    i0 = 200.0 * bell_centered(x, bbox, width=0.5) + 1.0
    y0 = np.concatenate((N * np.ones_like(i0) - i0, i0, np.zeros_like(i0)))

    def f(t, y):
        s, i, r = np.split(y, 3)
        tot = s + i + r
        return np.concatenate((-beta * s * i / tot, beta * s * i / tot - gamma * i, gamma * i))

    def df(t, y):
        s, i, r = np.split(y, 3)
        tot = s + i + r
        g = beta * s * i / tot  # infection term; d/ds, d/di, d/dr below
        gs = beta * i / tot - g / tot
        gi = beta * s / tot - g / tot
        gr = -g / tot
        D = np.diag
        Z = np.zeros_like
        return np.block([[D(-gs), D(-gi), D(-gr)],
                         [D(gs), D(gi - gamma), D(gr)],
                         [D(Z(s)), D(gamma * np.ones_like(s)), D(Z(s))]])

    return _problem(pts, kernel, diffusion_rates, "neumann", y0, t0, tmax, f=f, df=df,
                    n_int=n_int, n_bnd=n_bnd, nugget=nugget)


def lotka_volterra_1d(*, bbox=(0.0, 1.0), dx=None, num=None, t0=0.0, tmax=10.0, a=0.5, b=0.05, c=0.05, d=0.5,
                      diffusion_rates=(0.1, 0.1), kernel=None, n_int=3, n_bnd=3, nugget=0.0):
    """examples.py:181-248."""
    pts = mesh_1d(bbox, step=dx, num=num)
    kernel = kernel or SE()
    x = pts[:, 0]
    y0 = np.concatenate((5.0 * np.ones_like(x), 20.0 * np.exp(-(x ** 2))))

    def f(t, y):
        u, v = np.split(y, 2)
        return np.concatenate((a * u - b * u * v, c * u * v - d * v))

    def df(t, y):
        u, v = np.split(y, 2)
        D = np.diag
        return np.block([[D(a - b * v), D(-b * u)], [D(c * v), D(c * u - d)]])

    return _problem(pts, kernel, diffusion_rates, "neumann", y0, t0, tmax, f=f, df=df,
                    n_int=n_int, n_bnd=n_bnd, nugget=nugget)


def with_member(prob, diff_scale=1.0, y0=None):
    """An ensemble member of ``prob``: diffusivity multiplied by ``diff_scale`` (scales L
    and E_sqrtm, mixins.py:37-38) and/or a different initial condition."""
    out = SimpleNamespace(**vars(prob))
    out.L = prob.L * diff_scale
    out.E_sqrtm = prob.E_sqrtm * diff_scale
    if y0 is not None:
        out.y0 = np.asarray(y0, dtype=np.float64)
    return out
