"""Square-root covariance algebra on the GPU (API of src/pnmol/base/sqrt.py).

Inputs are CUDA float64 tensors (anything else is moved to the current CUDA device);
every function accepts an optional leading batch dimension, which replaces the reference's
``jax.vmap`` wrappers (sqrt.py:27-30).  The work is done by ``pnmol_b200_sqrt_propagate`` /
``pnmol_b200_sqrt_update`` (Householder QR with LAPACK ``dlarfg`` sign conventions).
"""
import torch

from .. import _lib


def _prep(x, device=None):
    if x is None:
        return None
    t = torch.as_tensor(x, dtype=torch.float64)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise _lib.PnmolB200Error("pnmol_b200.base.sqrt needs a CUDA device (no CPU fallback)")
        t = t.to(device or torch.device("cuda", torch.cuda.current_device()))
    return t.contiguous()


def propagate_cholesky_factor(S1, S2):
    """Cholesky factor of S1 S1^T + S2 S2^T (sqrt.py:9-12)."""
    S1 = _prep(S1)
    S2 = _prep(S2, S1.device)
    batched = S1.dim() == 3
    a = S1 if batched else S1[None]
    b = S2 if batched else S2[None]
    B, r, c1 = a.shape
    c2 = b.shape[2]
    out = torch.empty((B, r, min(r, c1 + c2)), dtype=torch.float64, device=a.device)
    lib = _lib.load()
    _lib.check(lib.pnmol_b200_sqrt_propagate(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), r, c1, c2, B, a.device.index,
                                             _lib.current_stream(a.device)))
    return out if batched else out[0]


def sqrtm_to_cholesky(St):
    """Lower factor from a 'right' square root St = S^T (sqrt.py:16-23)."""
    St = _prep(St)
    S = St.transpose(-1, -2).contiguous()
    empty = S.new_empty(S.shape[:-1] + (0,))
    batched = S.dim() == 3
    a = S if batched else S[None]
    B, r, c1 = a.shape
    out = torch.empty((B, r, min(r, c1)), dtype=torch.float64, device=a.device)
    lib = _lib.load()
    _lib.check(lib.pnmol_b200_sqrt_propagate(_lib.ptr(a), None, _lib.ptr(out), r, c1, 0, B, a.device.index,
                                             _lib.current_stream(a.device)))
    del empty
    return out if batched else out[0]


def _update(H, C, meascov):
    H = _prep(H)
    C = _prep(C, H.device)
    E = _prep(meascov, H.device)
    batched = H.dim() == 3
    h, c = (H, C) if batched else (H[None], C[None])
    e = None if E is None else (E if batched else E[None])
    B, m, D = h.shape
    C_out = torch.empty((B, D, D), dtype=torch.float64, device=h.device)
    K_out = torch.empty((B, D, m), dtype=torch.float64, device=h.device)
    S_out = torch.empty((B, m, m), dtype=torch.float64, device=h.device)
    lib = _lib.load()
    _lib.check(lib.pnmol_b200_sqrt_update(_lib.ptr(h), _lib.ptr(c), _lib.ptr(e), _lib.ptr(C_out), _lib.ptr(K_out),
                                          _lib.ptr(S_out), m, D, B, h.device.index, _lib.current_stream(h.device)))
    if batched:
        return C_out, K_out, S_out
    return C_out[0], K_out[0], S_out[0]


def update_sqrt(transition_matrix, cov_cholesky, meascov_sqrtm):
    """(posterior factor, Kalman gain, innovation factor) with measurement noise (sqrt.py:34-73)."""
    return _update(transition_matrix, cov_cholesky, meascov_sqrtm)


def update_sqrt_no_meascov(transition_matrix, cov_cholesky):
    """Noise-free update (sqrt.py:77-95)."""
    return _update(transition_matrix, cov_cholesky, None)


batched_propagate_cholesky_factor = propagate_cholesky_factor
batched_sqrtm_to_cholesky = sqrtm_to_cholesky
