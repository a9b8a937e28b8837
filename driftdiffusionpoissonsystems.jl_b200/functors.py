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


# --------------------------------------------------------------------------------------------
# Registered SPATIAL functors: coefficient closures (x,y)->value the library samples on the DEVICE
# (include/ddps.h: ddps_spatial_kind / ddps_coeff_functor) at the points where the reference evaluates
# them -- centroids (src:250-254, 463-467), edge midpoints (src:303-308, 341-343) -- so the one-time setup
# needs no host sampling and no upload.  Each is also an ordinary vectorised callable: user code, the host
# sampler and the CPU oracle evaluate it like the closure it replaces, and the device evaluates the SAME
# expression (no FMA contraction), so device-sampled arrays equal host-sampled ones bitwise.
# --------------------------------------------------------------------------------------------
SPATIAL_CONST, SPATIAL_AFFINE, SPATIAL_STEP_X, SPATIAL_STEP_Y, SPATIAL_BY_TAG = 1, 2, 3, 4, 5


class SpatialFunctor:
    kind = 0

    def params(self):
        raise NotImplementedError


class Const(SpatialFunctor):
    """(x,y) -> c"""
    kind = SPATIAL_CONST

    def __init__(self, c):
        self.c = float(c)

    def __call__(self, x, y):
        return self.c + 0.0 * np.asarray(x, dtype=np.float64)

    def params(self):
        return [self.c]


class Affine(SpatialFunctor):
    """(x,y) -> (a + bx*x) + by*y"""
    kind = SPATIAL_AFFINE

    def __init__(self, a, bx, by):
        self.a, self.bx, self.by = float(a), float(bx), float(by)

    def __call__(self, x, y):
        return (self.a + self.bx * np.asarray(x, dtype=np.float64)) + self.by * np.asarray(y, dtype=np.float64)

    def params(self):
        return [self.a, self.bx, self.by]


class StepX(SpatialFunctor):
    """(x,y) -> right if x > x0 else left   (abrupt pn doping profile)"""
    kind = SPATIAL_STEP_X

    def __init__(self, x0, left, right):
        self.x0, self.left, self.right = float(x0), float(left), float(right)

    def __call__(self, x, y):
        return np.where(np.asarray(x, dtype=np.float64) > self.x0, self.right, self.left)

    def params(self):
        return [self.x0, self.left, self.right]


class StepY(SpatialFunctor):
    """(x,y) -> above if y > y0 else below"""
    kind = SPATIAL_STEP_Y

    def __init__(self, y0, below, above):
        self.y0, self.below, self.above = float(y0), float(below), float(above)

    def __call__(self, x, y):
        return np.where(np.asarray(y, dtype=np.float64) > self.y0, self.above, self.below) + 0.0 * np.asarray(x, dtype=np.float64)

    def params(self):
        return [self.y0, self.below, self.above]
