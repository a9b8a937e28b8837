"""Registered u-dependent right-hand-side functors ("Functors").

The reference takes arbitrary closures ``f(x,y,u)`` and ``g(x,y,u)=df/du``
(``/root/reference/src/DriftDiffusionPoissonSystems.jl:895``).  Arbitrary host closures cannot run on
the device and there is no CPU fallback, so the u-dependence must come from this small registered
set.  Each functor is ALSO an ordinary callable ``(x,y,u)->value`` so that user code (and the CPU
oracle in the tests) can evaluate it like the closure it replaces.
"""
from __future__ import annotations

import numpy as np

# enum values shared with include/ddps.h (ddps_functor_kind)
FUNCTOR_SINH = 1
FUNCTOR_POLY = 2


class SinhRHS:
    """f(x,y,u) = h(x,y) + alpha*sinh(beta*u);  g = alpha*beta*cosh(beta*u).
    Covers ``test/runtests.jl:95-96`` with alpha=-1, beta=1."""

    kind = FUNCTOR_SINH

    def __init__(self, h, alpha=-1.0, beta=1.0):
        self.h, self.alpha, self.beta = h, float(alpha), float(beta)

    def __call__(self, x, y, u):
        return self.h(x, y) + self.alpha * np.sinh(self.beta * u)

    def derivative(self):
        return _SinhDerivative(self)


class _SinhDerivative:
    def __init__(self, parent):
        self.parent = parent

    def __call__(self, x, y, u):
        p = self.parent
        return p.alpha * p.beta * np.cosh(p.beta * u) + 0 * x


class PolyRHS:
    """f(x,y,u) = sum_k c_k(x,y) u^k, k=0..d;  g = sum_k k c_k(x,y) u^(k-1).
    ``coeffs`` is a list of closures (x,y)->c_k.  Covers the README example f=4xy u^2
    (``/root/reference/README.md:107-108``)."""

    kind = FUNCTOR_POLY
    MAX_DEGREE = 4

    def __init__(self, coeffs):
        if not 1 <= len(coeffs) <= self.MAX_DEGREE + 1:
            raise ValueError("PolyRHS supports degree 0..%d" % self.MAX_DEGREE)
        self.coeffs = list(coeffs)

    def __call__(self, x, y, u):
        out = 0.0
        for k, c in enumerate(self.coeffs):
            out = out + c(x, y) * u ** k
        return out

    def derivative(self):
        return _PolyDerivative(self)


class _PolyDerivative:
    def __init__(self, parent):
        self.parent = parent

    def __call__(self, x, y, u):
        out = 0.0 * x
        for k, c in enumerate(self.parent.coeffs):
            if k >= 1:
                out = out + k * c(x, y) * u ** (k - 1)
        return out
