"""NumPy emulation of the CUDA step (tests only).

Mirrors the device phases of pnmol-experiments_b200/csrc/ek1_device.cuh one to one --
same workspace layout, same support envelopes (taken from ``pnmol_b200_structure``), same
sparse H, same closed forms -- so that the *algorithm* the kernels implement can be
checked against the oracle without a GPU.  The workspace is NaN-poisoned: any read
outside what the build phases wrote shows up as NaN in the outputs.
"""
import ctypes

import numpy as np

from pnmol_b200 import _engine, _lib
from pnmol_b200.base import iwp as _iwp


def structure(kind, d, nu, nb, lcol, bcol, ncomp, dense=0):
    lib = _lib.load()
    n = nu + 1
    dd = 2 * d if kind >= 2 else d
    D, m = n * dd, d + nb
    te_p = np.zeros(D, np.int32); be_p = np.zeros(D, np.int32)
    te_u = np.zeros(m + D, np.int32); be_u = np.zeros(m + D, np.int32)
    _lib.check(lib.pnmol_b200_structure(kind, d, nu, nb, _lib.ptr(lcol), lcol.shape[1], _lib.ptr(bcol), bcol.shape[1],
                                        ncomp, dense, _lib.ptr(te_p), _lib.ptr(be_p), _lib.ptr(te_u), _lib.ptr(be_u)))
    return te_p, be_p, te_u, be_u


def householder_qr(W, nt, nbot, ncols, te, be):
    """W: (rows_alloc, >= ncols) array acting as the column-major workspace."""
    nrows = nt + nbot
    nref = min(nrows, ncols)
    for j in range(nref):
        if j < nt:
            e1 = min(te[j] if te is not None else nt - 1, nt - 1)
            e1 = max(e1, j)
            e2 = min(be[j] if be is not None else nrows - 1, nrows - 1)
            rows = list(range(j, e1 + 1)) + list(range(nt, e2 + 1))
        else:
            e1 = min(be[j] if be is not None else nrows - 1, nrows - 1)
            e1 = max(e1, j)
            rows = list(range(j, e1 + 1))
        rows = np.asarray(rows)
        v = W[rows, j].copy()
        ss = float(np.sum(v[1:] ** 2))
        alpha = v[0]
        if ss == 0.0 or np.isnan(ss) and False:
            continue
        nrm = np.sqrt(alpha * alpha + ss)
        beta = -np.copysign(nrm, alpha)
        tau = (beta - alpha) / beta
        v = v / (alpha - beta)
        v[0] = 1.0
        W[rows, j] = 0.0
        W[j, j] = beta
        if j + 1 < ncols:
            blk = W[np.ix_(rows, np.arange(j + 1, ncols))]
            w = tau * (v @ blk)
            W[np.ix_(rows, np.arange(j + 1, ncols))] = blk - np.outer(v, w)
    return W


NB, RPL, WARPS = 16, 16, 8


def _env(te, be, nt, nbot, j):
    nrows = nt + nbot
    et = min(te[j] if te is not None else nt - 1, nt - 1)
    eb = min(be[j] if be is not None else nrows - 1, nrows - 1)
    return et, eb


def panel_rows(nt, nbot, te, be, j0, jl):
    """Mirror of panel_rows in csrc/qr_blocked.cuh: compact row list of a panel."""
    if j0 < nt:
        jt = min(jl, nt - 1)
        e1 = max(_env(te, be, nt, nbot, jt)[0], jt)
        e2 = _env(te, be, nt, nbot, jl)[1]
        rows = list(range(j0, e1 + 1)) + list(range(nt, e2 + 1))
    else:
        e2 = max(_env(te, be, nt, nbot, jl)[1], jl)
        rows = list(range(j0, e2 + 1))
    return np.asarray(rows)


def householder_qr_blocked(W, nt, nbot, ncols, te, be, vld=512):
    """Mirror of householder_qr_blocked / qr_panel_step (csrc/qr_blocked.cuh): panels of NB columns, panel
    columns loaded with their own-envelope mask, trailing columns loaded unmasked on the panel's row list."""
    nrows = nt + nbot
    nref = min(nrows, ncols)
    j0 = 0
    while j0 < nref:
        nbk = min(NB, nref - j0)
        rows = panel_rows(nt, nbot, te, be, j0, j0 + nbk - 1)
        G = 4 if len(rows) <= 4 * RPL else 8 if len(rows) <= 8 * RPL else 16 if len(rows) <= 16 * RPL else 32
        cap = WARPS * (32 // G)
        if nbk > cap:
            nbk = cap
            rows = panel_rows(nt, nbot, te, be, j0, j0 + nbk - 1)
        assert len(rows) <= 32 * RPL and len(rows) <= vld, "fallback path not modelled"
        X = np.zeros((len(rows), nbk))
        masks = []
        for p in range(nbk):
            et, eb = _env(te, be, nt, nbot, j0 + p)
            mask = np.where(rows < nt, rows <= et, rows <= eb)
            masks.append(mask)
            X[mask, p] = W[rows[mask], j0 + p]
        V = np.zeros((len(rows), nbk))
        taus = np.zeros(nbk)
        for i in range(nbk):
            ss = float(np.sum(X[i + 1:, i] ** 2))
            al = X[i, i]
            if ss != 0.0:
                nrm = np.sqrt(al * al + ss)
                beta = -np.copysign(nrm, al)
                taus[i] = (beta - al) / beta
                V[i + 1:, i] = X[i + 1:, i] / (al - beta)
                V[i, i] = 1.0
                X[i:, i] = 0.0
                X[i, i] = beta
                w = taus[i] * (V[:, i] @ X[:, i + 1:])
                X[:, i + 1:] -= np.outer(V[:, i], w)
        for p in range(nbk):
            W[rows[masks[p]], j0 + p] = X[masks[p], p]
        if j0 + nbk < ncols:
            cols = np.arange(j0 + nbk, ncols)
            C = W[np.ix_(rows, cols)]
            for i in range(nbk):
                if taus[i] != 0.0:
                    C -= np.outer(V[:, i], taus[i] * (V[:, i] @ C))
            W[np.ix_(rows, cols)] = C
        j0 += nbk
    return W


def reaction_point(rid, prm, x):
    if rid == 1:
        g = prm[0]
        return np.array([g * x[0] * (1 - x[0])]), np.array([[g * (1 - 2 * x[0])]])
    if rid == 2:
        beta, gamma = prm[:2]
        s, i, r = x
        tot = s + i + r
        g = beta * s * i / tot
        gs, gi, gr = beta * i / tot - g / tot, beta * s / tot - g / tot, -g / tot
        return (np.array([-g, g - gamma * i, gamma * i]),
                np.array([[-gs, -gi, -gr], [gs, gi - gamma, gr], [0.0, gamma, 0.0]]))
    if rid == 3:
        a, b, c, dd = prm[:4]
        u, v = x
        return (np.array([a * u - b * u * v, c * u * v - dd * v]),
                np.array([[a - b * v, -b * u], [c * v, c * u - dd]]))
    raise ValueError(rid)


class Model:
    def __init__(self, pde, family, nu, gram_sqrtm, diff_scale=None, prior_scale=1.0, rparams=None, blocked=True):
        self.blocked = blocked
        self.latent = family == "latent"
        self.semil = bool(getattr(pde, "is_semilinear", False))
        self.kind = _engine.KINDS[(family, self.semil)]
        self.rid = pde.reaction.id if self.semil else 0
        self.prm = np.asarray(rparams if rparams is not None else (pde.reaction.params if self.semil else ()), float)
        self.lcol, self.lval = _engine.to_ell(pde.L)
        self.bcol, self.bval = _engine.to_ell(pde.B)
        self.ediag = np.diag(pde.E_sqrtm).copy()
        self.Rsq = np.asarray(pde.R_sqrtm, float)
        self.d, self.nb = pde.L.shape[0], pde.B.shape[0]
        self.nu, self.n = nu, nu + 1
        self.ncomp = getattr(pde, "num_components", 1)
        self.npts = self.d // self.ncomp
        self.dd = 2 * self.d if self.latent else self.d
        self.D, self.m = self.n * self.dd, self.d + self.nb
        self.ld = 2 * self.D
        self.A1d, self.LQ1d = _iwp.IntegratedWienerTransition(1, nu, np.eye(1)).preconditioned_discretize_1d
        self.Lk = np.asarray(gram_sqrtm, float)
        self.Kg = self.Lk @ self.Lk.T
        self.ds = np.ones(self.ncomp) if diff_scale is None else np.broadcast_to(np.asarray(diff_scale, float), (self.ncomp,))
        self.ps = prior_scale
        self.te_p, self.be_p, self.te_u, self.be_u = structure(self.kind, self.d, nu, self.nb, self.lcol, self.bcol, self.ncomp)
        self.te_pd = np.full(self.D, self.D - 1, np.int32)
        self.wh = max(self.lcol.shape[1] + (self.ncomp if self.semil else 0) + 1 + (1 if self.latent else 0), self.bcol.shape[1])
        self.W = np.full((self.ld, self.m + self.D), np.nan)

    # ---------------------------------------------------------------- phases
    def evaluate_ode(self, mp, p0s, p1s):
        n, d = self.n, self.d
        xat = p0s * mp[0::n][: self.dd]
        Hc = -np.ones((self.m, self.wh), np.int64)
        Hv = np.zeros((self.m, self.wh))
        z = np.zeros(self.m)
        for i in range(d):
            comp, pt = divmod(i, self.npts)
            acc, w = 0.0, 0
            for w in range(self.lcol.shape[1]):
                c = self.lcol[i, w]
                if c >= 0:
                    lv = self.ds[comp] * self.lval[i, w]
                    acc += lv * xat[c]
                    Hc[i, w], Hv[i, w] = c * n, -p0s * lv
            w = self.lcol.shape[1]
            shift = 0.0
            if self.semil:
                x = np.array([xat[c * self.npts + pt] for c in range(self.ncomp)])
                f, J = reaction_point(self.rid, self.prm, x)
                jx = 0.0
                for c in range(self.ncomp):
                    jx += J[comp, c] * x[c]
                    Hc[i, w], Hv[i, w] = (c * self.npts + pt) * n, -p0s * J[comp, c]
                    w += 1
                acc += jx
                shift = jx - f[comp]
            Hc[i, w], Hv[i, w] = i * n + 1, p1s
            w += 1
            hz = p1s * mp[i * n + 1] - acc
            if self.latent:
                Hc[i, w], Hv[i, w] = (d + i) * n, -p0s
                hz -= xat[d + i]
            z[i] = hz + shift
        for r in range(self.nb):
            acc = 0.0
            for w in range(self.bcol.shape[1]):
                c = self.bcol[r, w]
                if c >= 0:
                    acc += self.bval[r, w] * xat[c]
                    Hc[d + r, w], Hv[d + r, w] = c * n, p0s * self.bval[r, w]
            z[d + r] = acc
        return z, Hc, Hv

    def build_predict(self, Cl, pinv, te):
        n, D, nd, m = self.n, self.D, self.n * self.d, self.m
        for i in range(D):
            blk, ii = divmod(i, n)
            tend = te[i]
            acc = np.zeros(tend + 1)
            for s in range(n):
                acc += self.A1d[ii, s] * (pinv[s] * Cl[blk * n + s, : tend + 1])
            self.W[: tend + 1, m + i] = acc
            for k in range(i + 1):
                kb, kk = divmod(k, n)
                if i < nd:
                    v = (self.ps * self.Lk[blk, kb]) * self.LQ1d[ii, kk]
                else:
                    comp = (blk - self.d) // self.npts
                    eb = self.ds[comp] * self.ediag[blk - self.d]
                    v = eb * self.LQ1d[ii, kk] if kb == blk else 0.0
                self.W[D + k, m + i] = v

    def error_estimate(self, z, Hc, Hv, p1s, dt):
        n, d, m = self.n, self.d, self.m
        LQ = self.LQ1d
        q00, q01, q11 = LQ[0] @ LQ[0], LQ[0] @ LQ[1], LQ[1] @ LQ[1]
        ps2 = self.ps ** 2
        At = np.zeros((m, d))
        for r in range(m):
            for w in range(self.wh):
                c = Hc[r, w]
                if c >= 0 and c % n == 0 and c < n * d:
                    At[r, c // n] += Hv[r, w]
        K = ps2 * self.Kg
        F = At @ K
        It = np.vstack((np.eye(d), np.zeros((self.nb, d))))
        S = q00 * F @ At.T + q01 * p1s * (F @ It.T + It @ F.T) + q11 * p1s ** 2 * It @ K @ It.T
        E = np.zeros((m, m))
        for r in range(d):
            E[r, r] = self.ds[r // self.npts] * self.ediag[r]
        E[d:, d:] = self.Rsq
        S = S + E @ E.T
        diagS = np.diag(S).copy()
        Lc = np.linalg.cholesky(S)
        u = np.linalg.solve(Lc, z)
        sigma = np.sqrt(u @ u / m)
        return dt * (np.sqrt(diagS[:d]) * sigma)

    def meas_entry(self, emode, nugget, r, rp):
        v = 0.0
        if emode in ("step", "step+nugget"):
            if r < self.d:
                if rp == r:
                    v = self.ds[r // self.npts] * self.ediag[r]
            elif rp >= self.d:
                v = self.Rsq[r - self.d, rp - self.d]
        if emode in ("nugget", "step+nugget") and rp == r:
            v += nugget
        return v

    def update_stage(self, mp, z, Hc, Hv, mcur, emode, nugget, Rsrc, te, be, pv=None):
        D, m, n = self.D, self.m, self.n
        nbot = 0 if emode == "none" else mcur
        nrows = D + nbot
        off = m - mcur
        W = self.W
        for k in range(D):
            tend = min(te[mcur + k] if te is not None else D - 1, D - 1)
            if Rsrc is not None:
                W[: k + 1, m + k] = Rsrc[k, : k + 1]
            W[k + 1: tend + 1, m + k] = 0.0
            bend = min(be[mcur + k] if be is not None else nrows - 1, nrows - 1)
            W[D: bend + 1, m + k] = 0.0
        for r in range(mcur):
            tend = min(te[r] if te is not None else D - 1, D - 1)
            col = np.zeros(tend + 1)
            for w in range(self.wh):
                c = Hc[r, w]
                if c >= 0:
                    rows = np.arange(0, min(c, tend) + 1)
                    col[rows] += Hv[r, w] * W[rows, m + c]
            W[: tend + 1, off + r] = col
            bend = min(be[r] if be is not None else nrows - 1, nrows - 1)
            for i in range(D, bend + 1):
                W[i, off + r] = self.meas_entry(emode, nugget, r, i - D)
        sub = W[:, off:]
        (householder_qr_blocked if self.blocked else householder_qr)(sub, D, nbot, mcur + D, te, be)
        R1 = np.triu(sub[:mcur, :mcur])
        y = np.linalg.solve(R1.T, z[:mcur])
        x = np.linalg.solve(R1, z[:mcur])
        diff = x @ x / mcur
        R2 = sub[:mcur, mcur: mcur + D]
        m_new = mp - R2.T @ y
        chol = np.zeros((D, D))
        for r in range(D):
            for c in range(r + 1):
                if mcur + c < nrows:
                    chol[r, c] = sub[mcur + c, mcur + r]
        if pv is not None:
            scale = np.tile(pv, self.dd)
            m_new = scale * m_new
            chol = scale[:, None] * chol
        return m_new, chol, diff

    # ---------------------------------------------------------------- drivers
    def flat_to_mean(self, flat):
        return flat.reshape(self.dd, self.n).T.copy()

    def mean_to_flat(self, mean):
        return mean.T.reshape(-1).copy()

    def initialize(self, y0, prior_scale0=1.0):
        n, d, D, nd = self.n, self.d, self.D, self.n * self.d
        nugget = 1e-6 if self.latent else 1e-10
        C0 = np.zeros((D, D))
        for r in range(D):
            rb, ri = divmod(r, n)
            for c in range(r + 1):
                cb, ci = divmod(c, n)
                if ri == ci:
                    if r < nd:
                        C0[r, c] = prior_scale0 * (self.ps * self.Lk[rb, cb])
                    elif rb == cb:
                        C0[r, c] = prior_scale0 * (self.ds[(rb - d) // self.npts] * self.ediag[rb - d])
        Hc = -np.ones((self.m, self.wh), np.int64); Hv = np.zeros((self.m, self.wh))
        for i in range(d):
            Hc[i, 0], Hv[i, 0] = i * n, 1.0
        z = np.zeros(self.m); z[:d] = -np.asarray(y0)
        m1, C1, _ = self.update_stage(np.zeros(D), z, Hc, Hv, d, "nugget", nugget, C0, None, None)
        z, Hc, Hv = self.evaluate_ode(m1, 1.0, 1.0)
        m2, C2, _ = self.update_stage(m1, z, Hc, Hv, self.m, "nugget" if self.latent else "step+nugget", nugget, C1, None, None)
        return self.flat_to_mean(m2), C2

    def step(self, mean, chol, dt, dense=False):
        pv, pinv = _engine.nordsieck_raw(self.nu, dt)
        n, D = self.n, self.D
        m_in = self.mean_to_flat(mean)
        mp = np.zeros(D)
        for k in range(D):
            j, i = divmod(k, n)
            mp[k] = sum(self.A1d[i, s] * (pinv[s] * m_in[j * n + s]) for s in range(n))
        z, Hc, Hv = self.evaluate_ode(mp, pv[0], pv[1])
        te = self.te_pd if dense else self.te_p
        self.build_predict(chol, pinv, te)
        (householder_qr_blocked if self.blocked else householder_qr)(self.W[:, self.m:], D, D, D, te, self.be_p)
        err = None
        if not self.latent:
            err = self.error_estimate(z, Hc, Hv, pv[1], dt)
        m_new, C_new, diff = self.update_stage(mp, z, Hc, Hv, self.m, "none" if self.latent else "step", 0.0, None,
                                               self.te_u, self.be_u, pv=pv)
        return self.flat_to_mean(m_new), C_new, err, diff
