"""Probabilistic finite differences (host-side set-up, batched NumPy).

Same results as src/pnmol/discretize.py:12-201 for 1-D meshes: per mesh point, solve the
stencil's kernel Gram system for the weights (a row of ``L``) and keep the posterior
variance of the approximation (diagonal of ``E_sqrtm``, quirk Q3).  All stencils of one
size are solved in a single batched ``numpy.linalg.solve``.
"""
import numpy as np
import scipy.linalg

from . import diffops, kernels


def _gram_blocks(kernel, nbrs):
    # (P, s) -> (P, s, s) without forming the P*s x P*s Gram matrix
    P, s = nbrs.shape
    a = np.repeat(nbrs, s, axis=1).reshape(-1, 1)
    b = np.tile(nbrs, (1, s)).reshape(-1, 1)
    return kernel(a, b).reshape(P, s, s)


def fd_coefficients(x, neighbors, kernel, diffop=None, nugget_gram_matrix=0.0):
    """Weights and uncertainty of one stencil (discretize.py:177-201)."""
    diffop = diffop or diffops.laplace()
    x = np.atleast_1d(np.asarray(x, dtype=np.float64)).reshape(-1)[:1]
    nb = np.asarray(neighbors, dtype=np.float64).reshape(1, -1)
    w, u = _coefficients_blocks(kernel, diffop, x, nb, nugget_gram_matrix)
    return w[0], u[0]


def _coefficients_blocks(kernel, diffop, x, nbrs, nugget):
    first, second = {"laplace": ("dxx", "dxxyy"), "gradient": ("dx", "dxy")}[diffop.name]
    s = nbrs.shape[1]
    gram = _gram_blocks(kernel, nbrs) + nugget * np.eye(s)
    rhs = kernel.derivative(first, x[:, None], nbrs)
    top = kernel.derivative(second, x, x)
    if isinstance(kernel, kernels.Matern52):  # discretize.py:184-197
        r, sc = kernel.input_scale, kernel.output_scale
        rhs = np.where(np.isnan(rhs), r ** 2 * sc ** 2 * 2.5 / (1.0 - 2.5), rhs)
        top = np.where(np.isnan(top), sc ** 2 * r ** 4 * 3 * 2.5 ** 2 / (2.0 - 3 * 2.5 + 2.5 ** 2), top)
    w = np.linalg.solve(gram, rhs[..., None])[..., 0]
    return w, top - np.sum(w * rhs, axis=1)


def fd_probabilistic(diffop, mesh_spatial, kernel=None, stencil_size_interior=3, stencil_size_boundary=3,
                     nugget_gram_matrix=0.0):
    """Dense ``L`` and diagonal ``E_sqrtm`` (discretize.py:12-113)."""
    if mesh_spatial.dimension != 1:
        raise NotImplementedError("closed-form kernel derivatives are 1-D only")
    kernel = kernel or kernels.SquareExponential()
    N = len(mesh_spatial)
    L = np.zeros((N, N))
    E = np.zeros((N, N))
    for (pts, _, rows), size in ((mesh_spatial.boundary, stencil_size_boundary),
                                 (mesh_spatial.interior, stencil_size_interior)):
        if len(rows) == 0:
            continue
        nbrs, idx = mesh_spatial.neighbours(pts, size)
        w, unc = _coefficients_blocks(kernel, diffop, pts[:, 0], nbrs[..., 0], nugget_gram_matrix)
        L[rows[:, None], idx] = w
        E[rows, rows] = unc
    return L, E


def fd_probabilistic_neumann_1d(mesh_spatial, kernel=None, stencil_size=2, nugget_gram_matrix=0.0):
    """Normal-derivative rows at both ends of a 1-D mesh (discretize.py:116-174)."""
    if stencil_size != 2:
        raise NotImplementedError
    kernel = kernel or kernels.SquareExponential()
    x = mesh_spatial.points[:, 0]
    N = len(x)
    grad = diffops.gradient()
    w, unc = _coefficients_blocks(kernel, grad, x[[0, -1]], np.array([[x[0], x[1]], [x[-1], x[-2]]]),
                                  nugget_gram_matrix)
    B = np.eye(N)[[0, 1, N - 1, N - 2]]
    diffmatrix = scipy.linalg.block_diag(-w[0][None, :], w[1][None, :])
    return diffmatrix @ B, np.diag(unc)


# ------------------------------------------------------------------------- batched, on the device
_DIFFOP_ID = {"gradient": 0, "laplace": 1}


def fd_coefficients_batched(x, neighbors, kernel_list, diffop=None, nugget_gram_matrix=0.0, device=None):
    """``fd_coefficients`` (discretize.py:177-201) for every point ``x[p]`` with stencil ``neighbors[p]`` and every kernel
    of ``kernel_list`` in one launch (one thread per (kernel, point)).  Returns torch CUDA tensors
    ``weights [B, P, s]`` and ``uncertainties [B, P]``."""
    import torch

    from . import _lib

    diffop = diffop or diffops.laplace()
    kind, params, white = kernels._device_kernel_table(kernel_list)
    if white != 0.0:
        raise NotImplementedError("white noise has no derivative (the reference cannot discretise with it either)")
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    nb = np.asarray(neighbors, dtype=np.float64).reshape(len(x), -1)
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    dev = torch.device("cuda", idx if idx is not None else torch.cuda.current_device())
    lib = _lib.load()
    B, P, s = len(params), len(x), nb.shape[1]
    par, xd, nbd = (torch.as_tensor(a, device=dev) for a in (params, x, np.ascontiguousarray(nb)))
    w = torch.empty((B, P, s), dtype=torch.float64, device=dev)
    u = torch.empty((B, P), dtype=torch.float64, device=dev)
    _lib.check(lib.pnmol_b200_fd_coefficients(kind, _lib.ptr(par), B, _lib.ptr(xd), _lib.ptr(nbd), P, s, _DIFFOP_ID[diffop.name],
                                              float(nugget_gram_matrix), _lib.ptr(w), _lib.ptr(u), dev.index,
                                              _lib.current_stream(dev)))
    return w, u


def fd_probabilistic_batched(diffop, mesh_spatial, kernel_list, stencil_size_interior=3, stencil_size_boundary=3,
                             nugget_gram_matrix=0.0, device=None):
    """``fd_probabilistic`` (discretize.py:12-113) for a list of kernels (a sweep over kernel hyper-parameters): dense
    ``L [B, N, N]`` and ``E_sqrtm [B, N, N]`` as torch CUDA tensors, every stencil system solved on the device."""
    import torch

    if mesh_spatial.dimension != 1:
        raise NotImplementedError("closed-form kernel derivatives are 1-D only")
    N, B = len(mesh_spatial), len(kernel_list)
    L = E = None
    for (pts, _, rows), size in ((mesh_spatial.boundary, stencil_size_boundary), (mesh_spatial.interior, stencil_size_interior)):
        if len(rows) == 0:
            continue
        nbrs, idx = mesh_spatial.neighbours(pts, size)
        w, unc = fd_coefficients_batched(pts[:, 0], nbrs[..., 0], kernel_list, diffop, nugget_gram_matrix, device)
        if L is None:
            L = torch.zeros((B, N, N), dtype=torch.float64, device=w.device)
            E = torch.zeros((B, N, N), dtype=torch.float64, device=w.device)
        r = torch.as_tensor(np.asarray(rows), device=w.device)
        c = torch.as_tensor(np.asarray(idx), device=w.device)
        L[:, r[:, None], c] = w
        E[:, r, r] = unc
    return L, E
