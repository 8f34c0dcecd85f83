"""Lock-step batched Brent minimiser.

The reference optimises every branch length with an independent call of
scipy.optimize.minimize_scalar(method='brent', bracket=(xa, xb, xc), tol=tol)
(gtr.py:879-891).  Here all branches run the SAME state machine side by side:
each iteration proposes one trial point per still-active branch, evaluates all
of them in ONE device call (+ one allreduce when the patterns are sharded) and
updates the per-branch state with masks.  Per element the arithmetic is that of
scipy's `Brent.optimize` with a 3-point bracket (scipy/optimize/_optimize.py,
`class Brent`: constants _mintol = 1e-11, _cg = 0.3819660, maxiter = 500,
convergence |x - xmid| < 2 tol1 - (b - a)/2), so given the same function values
it takes the same iterates.  Because every decision is taken from the (reduced)
function values only, all ranks of a sharded run follow identical trajectories.
"""
import numpy as np

_MINTOL = 1.0e-11
_CG = 0.3819660


class BracketError(ValueError):
    """Same condition and message as scipy's ValueError for an invalid bracket."""


def brent_lockstep(func, xa, xb, xc, tol=1.48e-8, maxiter=500):
    """Minimise n independent scalar functions.

    func(active_idx, u) -> f : evaluates function `active_idx[k]` at `u[k]` (arrays).
    xa, xb, xc : arrays (n,), bracket with xa < xb < xc and f(xb) < f(xa), f(xb) < f(xc).
    Returns dict(x, fun, nit, nfev, success) of arrays.
    """
    xa = np.array(xa, dtype=float)
    xb = np.array(xb, dtype=float)
    xc = np.array(xc, dtype=float)
    n = xb.shape[0]
    swap = xa > xc
    xa[swap], xc[swap] = xc[swap], xa[swap]
    if not np.all((xa < xb) & (xb < xc)):
        raise BracketError('Bracketing values (xa, xb, xc) do not fulfill this requirement: (xa < xb) and (xb < xc)')
    everyone = np.arange(n)
    fa = np.asarray(func(everyone, xa), dtype=float)
    fb = np.asarray(func(everyone, xb), dtype=float)
    fc = np.asarray(func(everyone, xc), dtype=float)
    if not np.all((fb < fa) & (fb < fc)):
        raise BracketError('Bracketing values (xa, xb, xc) do not fulfill this requirement: (f(xb) < f(xa)) and (f(xb) < f(xc))')
    nfev = np.full(n, 3, dtype=np.int64)
    x = xb.copy(); w = xb.copy(); v = xb.copy()
    fx = fb.copy(); fw = fb.copy(); fv = fb.copy()
    a = xa.copy(); b = xc.copy()
    deltax = np.zeros(n)
    rat = np.zeros(n)
    nit = np.zeros(n, dtype=np.int64)
    active = np.ones(n, dtype=bool)
    with np.errstate(all='ignore'):
        while True:
            active &= nit < maxiter
            tol1 = tol * np.abs(x) + _MINTOL
            tol2 = 2.0 * tol1
            xmid = 0.5 * (a + b)
            active &= ~(np.abs(x - xmid) < (tol2 - 0.5 * (b - a)))
            if not active.any():
                break
            golden = np.abs(deltax) <= tol1
            # parabolic candidate
            tmp1 = (x - w) * (fx - fv)
            tmp2 = (x - v) * (fx - fw)
            p = (x - v) * tmp2 - (x - w) * tmp1
            tmp2 = 2.0 * (tmp2 - tmp1)
            p = np.where(tmp2 > 0.0, -p, p)
            tmp2 = np.abs(tmp2)
            ok = (~golden) & (p > tmp2 * (a - x)) & (p < tmp2 * (b - x)) & (np.abs(p) < np.abs(0.5 * tmp2 * deltax))
            # golden-section step (used when |deltax| <= tol1 or the parabola is rejected)
            gdelta = np.where(x >= xmid, a - x, b - x)
            grat = _CG * gdelta
            prat = p / tmp2
            u_par = x + prat
            near = ((u_par - a) < tol2) | ((b - u_par) < tol2)
            prat = np.where(near, np.where(xmid - x >= 0, tol1, -tol1), prat)
            new_deltax = np.where(ok, rat, gdelta)      # parabolic: deltax <- previous rat
            new_rat = np.where(ok, prat, grat)
            deltax = np.where(active, new_deltax, deltax)
            rat = np.where(active, new_rat, rat)
            u = np.where(np.abs(rat) < tol1, np.where(rat >= 0, x + tol1, x - tol1), x + rat)
            idx = np.nonzero(active)[0]
            fu_act = np.asarray(func(idx, u[idx]), dtype=float)
            fu = np.full(n, np.inf)
            fu[idx] = fu_act
            nfev[idx] += 1
            worse = active & (fu > fx)
            better = active & ~(fu > fx)
            # fu > fx: shrink the bracket around x, maybe update w, v
            a = np.where(worse & (u < x), u, a)
            b = np.where(worse & ~(u < x), u, b)
            c1 = worse & ((fu <= fw) | (w == x))
            c2 = worse & ~c1 & ((fu <= fv) | (v == x) | (v == w))
            v = np.where(c1, w, np.where(c2, u, v))
            fv = np.where(c1, fw, np.where(c2, fu, fv))
            w = np.where(c1, u, w)
            fw = np.where(c1, fu, fw)
            # fu <= fx: u becomes the new best point
            a = np.where(better & (u >= x), x, a)
            b = np.where(better & ~(u >= x), x, b)
            v = np.where(better, w, v); fv = np.where(better, fw, fv)
            w = np.where(better, x, w); fw = np.where(better, fx, fw)
            x = np.where(better, u, x); fx = np.where(better, fu, fx)
            nit[idx] += 1
    success = (nit < maxiter) & ~(np.isnan(x) | np.isnan(fx))
    return dict(x=x, fun=fx, nit=nit, nfev=nfev, success=success)
