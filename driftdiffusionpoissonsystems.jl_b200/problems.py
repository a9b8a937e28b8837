"""Problem definitions (closures + bddata) used by tests, golden generation and bench.py.

These are the reference's own test problems (``/root/reference/test/runtests.jl``, cited as
``test:N``) plus the pn-junction diode SURVEY.md section 8d defines for the large configs (the
reference ships none).  Everything here is plain closures exactly as a user of the reference would
write them; nothing imports the oracle or the CUDA library.
"""
from __future__ import annotations

import math

import numpy as np

from .functors import Const, SinhRHS, StepX


def ddpe1(c: float = 1.0):
    """test:14-35 -- all eight boundary segments Dirichlet, V=c(x+y), u=exp(-c(x+y)), v=exp(c(x+y)),
    rec_mode :zero, lambda=delta=mu=1, C=0, tol 1e-12."""
    tags = [1, 2, 3, 4, 5, 6, 7, 8]
    D = ["D"] * 8
    fV = lambda x, y: c * (x + y)
    fu = lambda x, y: np.exp(-c * (x + y))
    fv = lambda x, y: np.exp(c * (x + y))
    return dict(rec_mode="zero", Vbddata=[tags, D, [fV] * 8], ubddata=[tags, D, [fu] * 8],
                vbddata=[tags, D, [fv] * 8], lam=lambda x, y: 1.0 + 0 * x, delta=1.0, tau_p=0.0, tau_n=0.0,
                mu_p=lambda x, y: 1.0 + 0 * x, mu_n=lambda x, y: 1.0 + 0 * x, C=lambda x, y: 0.0 * x, tol=1e-12,
                exact=dict(V=fV, u=fu, v=fv))


def ddpe2(c: float = 1.0):
    """test:50-72 -- segments 1,3,5,7 Dirichlet, 2,4,6,8 Neumann, :shockley, tau_p=tau_n=1."""
    tags = [1, 2, 3, 4, 5, 6, 7, 8]
    kinds = ["D", "N", "D", "N", "D", "N", "D", "N"]
    zero = lambda x, y: 0.0 * x
    fV1 = lambda x, y: c + 0 * x
    fVl = lambda x, y: c * (x + y)
    fu = lambda x, y: math.exp(-c) + 0 * x
    fv = lambda x, y: math.exp(c) + 0 * x
    return dict(rec_mode="shockley",
                Vbddata=[tags, kinds, [fV1, zero, fVl, zero, fVl, zero, fVl, zero]],
                ubddata=[tags, kinds, [fu, zero] * 4], vbddata=[tags, kinds, [fv, zero] * 4],
                lam=lambda x, y: 1.0 + 0 * x, delta=1.0, tau_p=1.0, tau_n=1.0,
                mu_p=lambda x, y: 1.0 + 0 * x, mu_n=lambda x, y: 1.0 + 0 * x, C=lambda x, y: 0.0 * x, tol=1e-12)


def pn_junction(dirichlet_tags, all_tags, x_junction=0.0, C0=1.0, U=0.3, lam=1.0, delta=1.0, tau=1.0, mu=1.0,
                tol=1e-9, left_tags=None):
    """pn-junction diode of SURVEY.md section 8d config 2(ii)/4/5: doping C=+C0 for x>x_junction,
    -C0 otherwise; ohmic contacts (Dirichlet, thermal equilibrium + applied bias U on the right
    contact) on ``dirichlet_tags``, homogeneous Neumann elsewhere; :shockley recombination.
    Contact values follow from delta^2(v e^-V - u e^V) + C = 0 (src:341): V = +-asinh(C0/(2 delta^2)),
    u = exp(-U_applied), v = exp(U_applied)."""
    vbi = math.asinh(C0 / (2 * delta ** 2))
    kinds = ["D" if t in dirichlet_tags else "N" for t in all_tags]
    left = set(left_tags if left_tags is not None else [])

    def bias(x):
        return np.where(np.asarray(x) > x_junction, U, 0.0)

    fV = lambda x, y: np.where(np.asarray(x) > x_junction, vbi, -vbi) + bias(x)
    fu = lambda x, y: np.exp(-bias(x))
    fv = lambda x, y: np.exp(bias(x))
    zero = lambda x, y: 0.0 * np.asarray(x)
    pick = lambda f: [f if k == "D" else zero for k in kinds]
    return dict(rec_mode="shockley", Vbddata=[list(all_tags), kinds, pick(fV)],
                ubddata=[list(all_tags), kinds, pick(fu)], vbddata=[list(all_tags), kinds, pick(fv)],
                lam=Const(lam), delta=delta, tau_p=tau, tau_n=tau, mu_p=Const(mu), mu_n=Const(mu),
                C=StepX(x_junction, -C0, C0), tol=tol)   # registered spatial functors: ordinary closures that the library
                                                         # can also sample on the device (functors.py)


def semlin_test(tags4=True):
    """test:93-101 -- f = 4 e^{-r^2}(1-r^2) + sinh(e^{-r^2}) - sinh(u), g = -cosh(u), A = I,
    exact solution u = e^{-r^2} (README:186-196).  ``tags4``: 4-segment mesh (F1) with segments 1,3
    Dirichlet and 2,4 Neumann -/+2x e^{-r^2}; otherwise the 8-segment mesh (F2) with the same outward
    normal convention."""
    ex = lambda x, y: np.exp(-x ** 2 - y ** 2)
    h = lambda x, y: 4 * np.exp(-x ** 2 - y ** 2) * (1 - x ** 2 - y ** 2) + np.sinh(np.exp(-x ** 2 - y ** 2))
    rhs = SinhRHS(h, alpha=-1.0, beta=1.0)
    gR = lambda x, y: -2 * x * np.exp(-x ** 2 - y ** 2)
    gL = lambda x, y: 2 * x * np.exp(-x ** 2 - y ** 2)
    A = [lambda t: 1.0, lambda t: 0.0, lambda t: 0.0, lambda t: 1.0]
    if tags4:
        bddata = [[1, 2, 3, 4], ["D", "N", "D", "N"], [ex, gR, ex, gL]]
    else:
        bddata = [[1, 2, 3, 4, 5, 6, 7, 8], ["D", "N", "N", "N", "D", "N", "N", "N"],
                  [ex, gR, gR, gR, ex, gL, gL, gL]]
    return dict(A=A, bddata=bddata, f=rhs, g=rhs.derivative(), tol=1e-14, exact=ex)
