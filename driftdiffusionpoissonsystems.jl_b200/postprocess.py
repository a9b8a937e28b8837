"""Terminal-current post-processing of the reference, host side.

    aquire_boundary_separate_edges(mesh, Endpoints, bddata)      src:662-727
    get_gradients(uu, vv, ww, ar)                                src:734-748
    calculate_current(mesh, rec_mode, Endpoints, Vbddata, ubddata, vbddata, lambda, delta,
                      tau_p, tau_n, mu_p, mu_n, C, tol=1e-9) -> (I, V, u, v)       src:753-887

These run once after a solve (not per iteration), so they stay on the host; the Gummel loop inside
``calculate_current`` (src:776-783) is the GPU path (``solve_ddpe``).  The reference finds the triangle
behind every contact edge with a ``findfirst`` over ALL triangles (src:807,829,851,873: O(ne*nme), unusable
at 16M triangles); here an edge -> first-triangle map is built in one pass and gives the same triangle.
"""
from __future__ import annotations

import math

import numpy as np

from .api import _sample2, aquire_boundary, solve_ddpe
from .mesh import Mesh

# 3-point Gauss-Legendre rule on [-1,1] (the reference calls FastGaussQuadrature.gausslegendre(3), src:792)
GL3_NODES = np.array([-math.sqrt(3.0 / 5.0), 0.0, math.sqrt(3.0 / 5.0)])
GL3_WEIGHTS = np.array([5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0])


def _isapprox(a, b):
    """Julia's default isapprox: |a-b| <= sqrt(eps)*max(|a|,|b|) (exact equality when one side is 0)."""
    return np.abs(a - b) <= math.sqrt(np.finfo(float).eps) * np.maximum(np.abs(a), np.abs(b))


def aquire_boundary_separate_edges(mesh: Mesh, Endpoints, bddata):
    """src:662-727 -> (neu_nodes1, neu_nodes2, gD, n, edges_left, edges_bottom, edges_right, edges_top); the four
    edge arrays are 4 x k slices of ``mesh.edges`` holding the Dirichlet edges on each side of the rectangle."""
    E = np.asarray(Endpoints, dtype=float)
    n1, n2, gD, n = aquire_boundary(mesh, bddata)
    tags, kinds, _ = bddata
    sel = []
    for j, k in enumerate(kinds):
        if k == "D":
            sel.extend(np.nonzero(mesh.edges[3] == tags[j])[0].tolist())       # src:690-691
    sel = np.array(sel, dtype=np.int64)
    a, b = mesh.edges[0, sel] - 1, mesh.edges[1, sel] - 1
    xpar = _isapprox(mesh.nodes[1, a] - mesh.nodes[1, b], 0.0)                   # src:701
    ypar = _isapprox(mesh.nodes[0, a] - mesh.nodes[0, b], 0.0)                   # src:702
    if int(xpar.sum()) + int(ypar.sum()) != sel.size:                           # src:704
        raise AssertionError("Dirichlet edges must be parallel to the axes")
    left = ypar & _isapprox(E[0, 0], mesh.nodes[0, a])                           # src:708
    bottom = xpar & _isapprox(E[0, 1], mesh.nodes[1, a])                         # src:709
    right = ypar & _isapprox(E[2, 0], mesh.nodes[0, a])                          # src:710
    top = xpar & _isapprox(E[2, 1], mesh.nodes[1, a])                            # src:711
    if int(left.sum() + bottom.sum() + right.sum() + top.sum()) != sel.size:    # src:713
        raise AssertionError("Dirichlet edges must lie on the sides of the rectangle given by Endpoints")
    pick = lambda m: mesh.edges[:, sel[m]]
    return n1, n2, gD, n, pick(left), pick(bottom), pick(right), pick(top)


def get_gradients(uu, vv, ww, ar):
    """src:734-748: x- and y-derivatives of the three P1 basis functions of every triangle (nme x 3 each)."""
    gradx = np.stack([0.5 * uu[1] / ar, 0.5 * vv[1] / ar, 0.5 * ww[1] / ar], axis=1)
    grady = np.stack([-0.5 * uu[0] / ar, -0.5 * vv[0] / ar, -0.5 * ww[0] / ar], axis=1)
    return gradx, grady


def _geometry(mesh: Mesh):
    a = mesh.elements[0:3] - 1
    q1, q2, q3 = mesh.nodes[:, a[0]], mesh.nodes[:, a[1]], mesh.nodes[:, a[2]]
    uu, vv, ww = q2 - q3, q3 - q1, q1 - q2                                       # src:243-245
    return uu, vv, ww, (uu[0] * vv[1] - uu[1] * vv[0]) / 2                      # src:246


def _first_triangle_of_edges(mesh: Mesh):
    """(min node, max node) -> smallest index of a triangle containing both (= the reference's findfirst)."""
    t = mesh.elements[0:3].T
    first = {}
    for e in range(t.shape[0] - 1, -1, -1):                                      # descending: the smallest index wins
        a, b, c = int(t[e, 0]), int(t[e, 1]), int(t[e, 2])
        first[(min(a, b), max(a, b))] = e
        first[(min(b, c), max(b, c))] = e
        first[(min(a, c), max(a, c))] = e
    return first


def current_from_fields(mesh: Mesh, Endpoints, Vbddata, mu_p, mu_n, u, v):
    """The flux integral of src:786-886 for given Slotboom fields: 3-point Gauss-Legendre along every Dirichlet
    edge of the four sides, J = mu_n e^{V} grad u - mu_p e^{-V} grad v with V taken from the boundary closure."""
    E = np.asarray(Endpoints, dtype=float)
    _, _, _, _, e_left, e_bottom, e_right, e_top = aquire_boundary_separate_edges(mesh, E, Vbddata)
    uu, vv, ww, ar = _geometry(mesh)
    gradx, grady = get_gradients(uu, vv, ww, ar)                                 # src:770
    first = _first_triangle_of_edges(mesh)
    u = np.asarray(u).reshape(-1)
    v = np.asarray(v).reshape(-1)
    total = 0.0
    # (edges, fixed coordinate, integrate along x?, gradient table, sign)  src:800-884
    sides = ((e_left, E[0, 0], False, gradx, -1.0), (e_bottom, E[0, 1], True, grady, -1.0),
             (e_right, E[2, 0], False, gradx, 1.0), (e_top, E[2, 1], True, grady, 1.0))
    for edges, fixed, along_x, gtab, sign in sides:
        for i in range(edges.shape[1]):
            na, nb, tag = int(edges[0, i]), int(edges[1, i]), int(edges[3, i])
            Vb = Vbddata[2][tag - 1]                                             # src:801 (column index = geometric tag)
            coord = mesh.nodes[0 if along_x else 1, [na - 1, nb - 1]]
            lo, hi = float(coord.min()), float(coord.max())                      # src:804-805
            ind = first[(min(na, nb), max(na, nb))]                              # src:807
            nodes3 = mesh.elements[0:3, ind] - 1
            gradu = float(np.dot(u[nodes3], gtab[ind]))                          # src:811
            gradv = float(np.dot(v[nodes3], gtab[ind]))
            s = (GL3_NODES + 1) * (hi - lo) / 2 + lo                             # src:798
            xs, ys = (s, np.full(3, fixed)) if along_x else (np.full(3, fixed), s)
            eV = np.exp(_sample2(Vb, xs, ys))
            Ip = _sample2(mu_n, xs, ys) * eV * gradu * (hi - lo) / 2             # src:802,814
            In = -_sample2(mu_p, xs, ys) / eV * gradv * (hi - lo) / 2            # src:803,815
            total += sign * float(np.dot(GL3_WEIGHTS, Ip + In))                  # src:817,861
    return total


def calculate_current(mesh, rec_mode, Endpoints, Vbddata, ubddata, vbddata, lam, delta, tau_p, tau_n, mu_p, mu_n, C_,
                      tol=1e-9, session=None, verbose=True):
    """Drop-in for ``calculate_current`` (src:753-887): Gummel loop on the GPU, flux integrals on the host.
    Returns ``(I, V, u, v)``."""
    V, u, v = solve_ddpe(mesh, rec_mode, Vbddata, ubddata, vbddata, lam, delta, tau_p, tau_n, mu_p, mu_n, C_, tol,
                         session=session, verbose=verbose)
    return current_from_fields(mesh, Endpoints, Vbddata, mu_p, mu_n, u, v), V, u, v
