"""CPU oracle for the DriftDiffusionPoissonSystems.jl hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy *restatement* (not a copy) of the reference algorithm in
``/root/reference/src/DriftDiffusionPoissonSystems.jl`` (cited below as ``src:N``).  It exists to
check the CUDA path; nothing in the product package imports it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it.

PARITY PINNING STATUS: **parity unpinned by the reference itself.**  The reference ships no golden
vectors, no ``@test`` and never calls its own test functions (``test/runtests.jl:1-112``); Julia is not
installed in this image so the reference cannot be executed here.  The oracle is therefore pinned
against (a) the analytic solutions the reference's test scripts print errors against
(``test/runtests.jl:19-21,44-47,95-110``) and (b) a dense-LU mode that follows the reference *as
written* (``full(...)`` + ``\\``) cross-checked against the sparse mode.  See tests/test_oracle.py.

Shims applied to the literal reference:
  S-a  ``J0 = full(sparse(...))`` (src:391,584,1062) is kept sparse in ``linsolve='sparse'`` mode.
  S-b  ``ar = ar'`` at src:247 makes ``./ar4`` a Vector./RowVector broadcast; the intended element-wise
       quotient (the form uvsolve uses, src:474-479) is what is restated.
  S-c  The Neumann branch of ``aquire_boundary`` appends to undefined ``Irhs/Krhs`` (src:192-194);
       those three lines are dropped (ddpe Neumann data are homogeneous anyway, src:356-358).

All node / element indices inside this module are 1-based in the ``Mesh`` container exactly like the
reference (src:9-18) and converted to 0-based at the point of use.
"""
from __future__ import annotations

import dataclasses
import time
from typing import Callable, Sequence

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


# --------------------------------------------------------------------------------------------
# Mesh container and readers (src:9-18, src:95-154)
# --------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Mesh:
    """src:9-18.  nodes 2 x nq float64; edges 4 x ne int64 [n1;n2;phys;geom]; elements 5 x nme
    int64 [n1;n2;n3;phys;geom]; node indices are 1-based."""

    nodes: np.ndarray
    edges: np.ndarray
    elements: np.ndarray


def read_mesh(path: str) -> Mesh:
    """gmsh MSH 2.2 ASCII reader, restating src:95-154 (only line (type 1) and triangle (type 2)
    elements with exactly two tags are accepted; trailing sections are ignored with a warning)."""
    with open(path, "r") as fh:
        lines = [ln.split() for ln in fh.read().splitlines()]
    i = 0
    assert lines[i] == ["$MeshFormat"]                      # src:100
    i += 1
    assert float(lines[i][0]) == 2.2                        # src:103
    i += 1
    assert lines[i] == ["$EndMeshFormat"]                   # src:105
    i += 1
    if lines[i] == ["$PhysicalNames"]:                      # src:108-112
        nr_names = int(lines[i + 1][0])
        i += nr_names + 3
    assert lines[i] == ["$Nodes"]                           # src:115
    nq = int(lines[i + 1][0])
    rows = np.array([[float(t) for t in ln] for ln in lines[i + 2:i + 2 + nq]])
    assert np.all(rows[:, 3] == 0.0)                        # src:119 planar mesh
    nodes = np.ascontiguousarray(rows[:, 1:3].T)            # src:120
    i += nq + 2
    assert lines[i] == ["$EndNodes"]                        # src:122
    i += 1
    assert lines[i] == ["$Elements"]                        # src:126
    nall = int(lines[i + 1][0])
    recs = lines[i + 2:i + 2 + nall]
    edges, elements = [], []
    for r in recs:
        etype, ntags = int(r[1]), int(r[2])
        if etype == 1:                                      # src:129-136
            assert ntags == 2
            edges.append([int(r[5]), int(r[6]), int(r[3]), int(r[4])])
        elif etype == 2:                                    # src:132-138
            assert ntags == 2
            elements.append([int(r[5]), int(r[6]), int(r[7]), int(r[3]), int(r[4])])
        else:                                               # src:140-142
            raise ValueError("Unknown type of elements used! Please only use edges and triangular elements!")
    i += nall + 2
    assert lines[i] == ["$EndElements"]                     # src:145
    return Mesh(nodes, np.array(edges, dtype=np.int64).T.copy(), np.array(elements, dtype=np.int64).T.copy())


def structured_mesh(nx: int, ny: int, x0=0.0, x1=1.0, y0=0.0, y1=1.0) -> Mesh:
    """Synthetic structured triangulation in the reference's Mesh layout (SURVEY.md section 8d,
    config 3): node (i,j) -> id j*(nx+1)+i (+1), each cell split along (i,j)-(i+1,j+1), both
    triangles CCW: (n00,n10,n11),(n00,n11,n01); boundary tags 1 bottom / 2 right / 3 top / 4 left."""
    xs = np.linspace(x0, x1, nx + 1)
    ys = np.linspace(y0, y1, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    nodes = np.vstack([X.ravel(), Y.ravel()])
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    n00 = (jj * (nx + 1) + ii).ravel() + 1
    n10, n01, n11 = n00 + 1, n00 + nx + 1, n00 + nx + 2
    el = np.empty((5, 2 * nx * ny), dtype=np.int64)
    el[0, 0::2], el[1, 0::2], el[2, 0::2] = n00, n10, n11
    el[0, 1::2], el[1, 1::2], el[2, 1::2] = n00, n11, n01
    el[3:, :] = 2
    i = np.arange(nx)
    j = np.arange(ny)
    bottom = np.vstack([i + 1, i + 2, np.full(nx, 1), np.full(nx, 1)])
    right = np.vstack([j * (nx + 1) + nx + 1, (j + 1) * (nx + 1) + nx + 1, np.full(ny, 1), np.full(ny, 2)])
    top = np.vstack([ny * (nx + 1) + nx + 1 - i, ny * (nx + 1) + nx - i, np.full(nx, 1), np.full(nx, 3)])
    left = np.vstack([(ny - j) * (nx + 1) + 1, (ny - j - 1) * (nx + 1) + 1, np.full(ny, 1), np.full(ny, 4)])
    edges = np.hstack([bottom, right, top, left]).astype(np.int64)
    return Mesh(nodes, edges, el)


# --------------------------------------------------------------------------------------------
# bddata handling (src:162-218 and its inline copies src:532-544, src:965-1013)
# --------------------------------------------------------------------------------------------
def _bd_columns(bddata):
    """bddata is the reference's 3 x n heterogeneous array (README:87-92): row 1 geometric tag,
    row 2 'D'/'N', row 3 closure (x,y)->value.  Accepts a list of 3 rows."""
    tags, kinds, fcts = bddata
    if not (len(tags) == len(kinds) == len(fcts)):
        raise ValueError("bddata must be 3 x n")
    neu = [j for j, k in enumerate(kinds) if k == "N"]      # src:166
    diri = [j for j, k in enumerate(kinds) if k == "D"]     # src:167
    if len(neu) + len(diri) != len(tags):                   # src:169-171
        raise ValueError("Could not understand boundary file. Not all parts of the boundary have BC assigned.")
    return tags, fcts, neu, diri


def _unique_in_order(a):
    _, first = np.unique(a, return_index=True)
    return a[np.sort(first)]


def dirichlet_data(mesh: Mesh, bddata):
    """Dirichlet node list ``n`` (1-based, order of first appearance) and values ``gD``
    (src:199-211; same code at src:536-544 and src:1001-1013)."""
    tags, fcts, _, diri = _bd_columns(bddata)
    diri_nodes: list = []
    diri_values: list = []
    for j in diri:
        idx = np.nonzero(mesh.edges[3, :] == tags[j])[0]                      # src:202
        seg = _unique_in_order(mesh.edges[0:2, idx].T.ravel())                 # src:203 (vec of 2 x k, column-major)
        diri_nodes.extend(int(s) for s in seg)
        if len(diri_nodes) < len(idx) + 1:
            raise ValueError("Dirichlet segment has fewer nodes than edges+1 (reference: BoundsError at src:204)")
        tail = diri_nodes[len(diri_nodes) - len(idx) - 1:]                     # last k+1 entries, src:204
        vals = [float(fcts[j](mesh.nodes[0, nd - 1], mesh.nodes[1, nd - 1])) for nd in tail]   # src:204
        diri_values.extend(vals)
    if len(diri_nodes) != len(diri_values):
        raise ValueError("Dirichlet segment is not a single open polyline (reference would throw DimensionMismatch at src:206)")
    seen = set()
    n_out, g_out = [], []
    for nd, val in zip(diri_nodes, diri_values):                               # src:206 unique rows, first occurrence
        if (nd, val) not in seen:
            seen.add((nd, val))
            n_out.append(nd)
            g_out.append(val)
    return np.array(n_out, dtype=np.int64), np.array(g_out, dtype=np.float64)


def neumann_data(mesh: Mesh, bddata):
    """Neumann edge end-node lists and the per-edge load len(E)/2*gN(edge midpoint)
    (src:976-991; the ddpe copy src:174-189 computes the same lists but its data are unused)."""
    tags, fcts, neu, _ = _bd_columns(bddata)
    n1: list = []
    n2: list = []
    fct: list = []
    for j in neu:
        idx = np.nonzero(mesh.edges[3, :] == tags[j])[0]                      # src:981
        a, b = mesh.edges[0, idx], mesh.edges[1, idx]
        n1.extend(a.tolist())
        n2.extend(b.tolist())
        mx = (mesh.nodes[0, a - 1] + mesh.nodes[0, b - 1]) / 2                 # src:984
        my = (mesh.nodes[1, a - 1] + mesh.nodes[1, b - 1]) / 2
        fct.extend(float(fcts[j](x, y)) for x, y in zip(mx, my))
    n1 = np.array(n1, dtype=np.int64)
    n2 = np.array(n2, dtype=np.int64)
    fct = np.array(fct, dtype=np.float64)
    if len(n1):
        length = np.sqrt((mesh.nodes[0, n1 - 1] - mesh.nodes[0, n2 - 1]) ** 2
                         + (mesh.nodes[1, n1 - 1] - mesh.nodes[1, n2 - 1]) ** 2)
        load = length * fct / 2                                               # src:990
    else:
        load = np.zeros(0)
    return n1, n2, load


def aquire_boundary(mesh: Mesh, bddata):
    """src:162-218 with shim S-c.  Returns (neu_nodes1, neu_nodes2, gD, n)."""
    n1, n2, _ = neumann_data(mesh, bddata)
    n, gD = dirichlet_data(mesh, bddata)
    return n1, n2, gD, n


# --------------------------------------------------------------------------------------------
# Element kernels (K1-K4 of SURVEY.md section 2.1)
# --------------------------------------------------------------------------------------------
def _geometry(mesh: Mesh):
    """src:235-247.  Corner ids (0-based), edge vectors uu=q2-q3, vv=q3-q1, ww=q1-q2 (2 x nme) and the
    SIGNED area ar=(uu_x*vv_y-uu_y*vv_x)/2."""
    a1, a2, a3 = (mesh.elements[k, :] - 1 for k in range(3))
    q1, q2, q3 = mesh.nodes[:, a1], mesh.nodes[:, a2], mesh.nodes[:, a3]
    uu, vv, ww = q2 - q3, q3 - q1, q1 - q2
    ar = (uu[0] * vv[1] - uu[1] * vv[0]) / 2
    return a1, a2, a3, uu, vv, ww, ar


def _stiffness_triplets(uu, vv, ww, ar, phx, phy):
    """src:257-269 / 471-483 / 924-936: K_ab = (e_ax*phx*e_bx + e_ay*phy*e_by)/|4 ar|, the 9 local
    entries in the order I=[1,2,3,1,2,3,1,2,3], J=[1,1,1,2,2,2,3,3,3]."""
    ar4 = np.abs(ar * 4)
    e = (uu, vv, ww)

    def k(a, b):
        return (e[a][0] * phx * e[b][0] + e[a][1] * phy * e[b][1]) / ar4

    Kg = np.empty((9, ar.size))
    Kg[0], Kg[1], Kg[2] = k(0, 0), k(1, 0), k(2, 0)
    Kg[4], Kg[5], Kg[8] = k(1, 1), k(2, 1), k(2, 2)
    Kg[[3, 6, 7]] = Kg[[1, 2, 5]]
    return Kg


def _wmass_triplets(W1, W2, W3):
    """src:380-386 / 573-579 / 1051-1057 (weighted mass with nodal weights W_k = g_k*ar/30)."""
    WKg = np.empty((9, W1.size))
    WKg[0] = 3 * W1 + W2 + W3
    WKg[1] = W1 + W2 + W3 / 2
    WKg[2] = W1 + W2 / 2 + W3
    WKg[4] = W1 + 3 * W2 + W3
    WKg[5] = W1 / 2 + W2 + W3
    WKg[8] = W1 + W2 + 3 * W3
    WKg[[3, 6, 7]] = WKg[[1, 2, 5]]
    return WKg


def _local_ij(a1, a2, a3):
    a = np.vstack([a1, a2, a3])
    Ig = a[[0, 1, 2, 0, 1, 2, 0, 1, 2], :]   # src:267
    Jg = a[[0, 0, 0, 1, 1, 1, 2, 2, 2], :]   # src:268
    return Ig.T.ravel(), Jg.T.ravel()         # element-major, like vec() of a 9 x nme matrix


def _sparse(I, J, V, nq):
    """Julia ``sparse(I,J,V,m,n)``: duplicates summed, explicit zeros kept (src:284 etc.)."""
    return sp.coo_matrix((V, (I, J)), shape=(nq, nq)).tocsr()


def _eliminate(Ig, Jg, Kg, n0, nq):
    """src:271-284 (K6): zero the triplets whose row or column is a Dirichlet node, append (n,n,1);
    bdM collects the removed triplets (row hits then column hits, D x D block double counted)."""
    isD = np.zeros(nq, dtype=bool)
    isD[n0] = True
    rows = isD[Ig]
    cols = isD[Jg]
    bdM = _sparse(np.concatenate([Ig[rows], Ig[cols]]), np.concatenate([Jg[rows], Jg[cols]]),
                  np.concatenate([Kg[rows], Kg[cols]]), nq)
    Kg = Kg.copy()
    Kg[rows] = 0
    Kg[cols] = 0
    M = _sparse(np.concatenate([Ig, n0]), np.concatenate([Jg, n0]), np.concatenate([Kg, np.ones(len(n0))]), nq)
    return M, bdM


def _load_vector(a1, a2, a3, fa, fb, fc, ar, nq, extra_idx=None, extra_val=None):
    """src:345-361 (K4): lumped edge-midpoint rule, node a1 <- f(m12)|ar|/3, a2 <- f(m23)|ar|/3,
    a3 <- f(m31)|ar|/3; optional Neumann triplets appended (src:994-996)."""
    w = np.abs(ar) / 3
    idx = np.concatenate([a1, a2, a3])
    val = np.concatenate([fa * w, fb * w, fc * w])
    if extra_idx is not None and len(extra_idx):
        idx = np.concatenate([idx, extra_idx])
        val = np.concatenate([val, extra_val])
    return np.bincount(idx, weights=val, minlength=nq)


def _solve(A, rhs, linsolve):
    if linsolve == "dense":          # the reference as written: full() then dense LU (src:391-394)
        return np.linalg.solve(A.toarray(), rhs)
    return spla.splu(A.tocsc()).solve(rhs)


# --------------------------------------------------------------------------------------------
# MV_assemble (src:229-287)
# --------------------------------------------------------------------------------------------
def MV_assemble(mesh: Mesh, n, gD, lam: Callable):
    nq = mesh.nodes.shape[1]
    a1, a2, a3, uu, vv, ww, ar = _geometry(mesh)
    cx = (mesh.nodes[0, a1] + mesh.nodes[0, a2] + mesh.nodes[0, a3]) / 3      # src:250
    cy = (mesh.nodes[1, a1] + mesh.nodes[1, a2] + mesh.nodes[1, a3]) / 3
    ph = _vec2(lam, cx, cy) ** 2                                               # src:254
    Kg = _stiffness_triplets(uu, vv, ww, ar, ph, ph).T.ravel()                 # src:260-269
    Ig, Jg = _local_ij(a1, a2, a3)
    M, bdM = _eliminate(Ig, Jg, Kg, np.asarray(n) - 1, nq)
    return M, bdM, ar, uu, vv, ww


def _vec2(fn, x, y):
    """Broadcast a scalar closure (x,y)->value like Julia's ``map``/dot-call; tolerates closures
    that return a constant."""
    try:
        out = np.asarray(fn(x, y), dtype=np.float64)
        if out.shape == np.shape(x):
            return out
        if out.shape == ():
            return np.full(np.shape(x), float(out))
    except Exception:
        pass
    return np.array([float(fn(a, b)) for a, b in zip(np.ravel(x), np.ravel(y))]).reshape(np.shape(x))


def _vec3(fn, x, y, u):
    try:
        out = np.asarray(fn(x, y, u), dtype=np.float64)
        if out.shape == np.shape(x):
            return out
        if out.shape == ():
            return np.full(np.shape(x), float(out))
    except Exception:
        pass
    return np.array([float(fn(a, b, c)) for a, b, c in zip(np.ravel(x), np.ravel(y), np.ravel(u))]).reshape(np.shape(x))


# --------------------------------------------------------------------------------------------
# Vsolve (src:292-425)
# --------------------------------------------------------------------------------------------
def _damped_update(V, dx, res, M, rhs, damping):
    """NOT in the reference (its Newton update is undamped, src:394-395,1063-1064): restatement of the library's opt-in
    ``opts.newton_damping`` so that the tests can check it -- halve the step (at most
    ``damping`` times) until the residual max-norm decreases.  damping = 0 is the reference."""
    t = 1.0
    Vn = V - dx
    bn = rhs(Vn)
    k = 0
    while k < damping and not np.max(np.abs(M @ Vn - bn)) < res:
        t /= 2
        Vn = V - t * dx
        bn = rhs(Vn)
        k += 1
    return Vn, bn


def Vsolve(mesh, u, v, delta, M, bdM, ar, n, gD, C, tol, linsolve="sparse", trace=None, max_newton=200, damping=0):
    nq = mesh.nodes.shape[1]
    a1, a2, a3 = (mesh.elements[k, :] - 1 for k in range(3))
    n0 = np.asarray(n) - 1
    X, Y = mesh.nodes
    hax, hay = (X[a1] + X[a2]) / 2, (Y[a1] + Y[a2]) / 2                         # src:303-308
    hbx, hby = (X[a2] + X[a3]) / 2, (Y[a2] + Y[a3]) / 2
    hcx, hcy = (X[a3] + X[a1]) / 2, (Y[a3] + Y[a1]) / 2
    Ca, Cb, Cc = _vec2(C, hax, hay), _vec2(C, hbx, hby), _vec2(C, hcx, hcy)
    u1, u2, u3 = u[a1], u[a2], u[a3]
    v1, v2, v3 = v[a1], v[a2], v[a3]
    ua, ub, uc = (u1 + u2) / 2, (u2 + u3) / 2, (u3 + u1) / 2                    # src:332-338
    va, vb, vc = (v1 + v2) / 2, (v2 + v3) / 2, (v3 + v1) / 2
    lift = bdM[:, n0] @ gD                                                      # src:362

    def rhs(V):
        V1, V2, V3 = V[a1], V[a2], V[a3]
        Sa, Sb, Sc = (V1 + V2) / 2, (V2 + V3) / 2, (V3 + V1) / 2                # src:318-320
        fa = delta ** 2 * (np.exp(-Sa) * va - np.exp(Sa) * ua) + Ca             # src:341-343
        fb = delta ** 2 * (np.exp(-Sb) * vb - np.exp(Sb) * ub) + Cb
        fc = delta ** 2 * (np.exp(-Sc) * vc - np.exp(Sc) * uc) + Cc
        b = _load_vector(a1, a2, a3, fa, fb, fc, ar, nq)                        # src:345-361
        b = b - lift
        b[n0] = gD                                                              # src:363
        return b

    V = np.zeros(nq)                                                            # src:311
    b = rhs(V)
    Ig, Jg = _local_ij(a1, a2, a3)
    newton = 0
    res_hist = []
    while True:
        res = np.max(np.abs(M @ V - b))                                         # src:368
        res_hist.append(res)
        if not res > tol or newton >= max_newton:
            break
        V1, V2, V3 = V[a1], V[a2], V[a3]
        W1 = -delta ** 2 * (u1 * np.exp(V1) + v1 * np.exp(-V1)) * ar / 30       # src:375-377 (signed ar)
        W2 = -delta ** 2 * (u2 * np.exp(V2) + v2 * np.exp(-V2)) * ar / 30
        W3 = -delta ** 2 * (u3 * np.exp(V3) + v3 * np.exp(-V3)) * ar / 30
        J0 = _sparse(Ig, Jg, _wmass_triplets(W1, W2, W3).T.ravel(), nq)         # src:379-391
        A = M - J0
        if trace is not None and "V_A_first" not in trace:
            trace["V_A_first"] = A.copy()
            trace["V_b_first"] = b.copy()
        V, b = _damped_update(V, _solve(A, M @ V - b, linsolve), res, M, rhs, damping)   # src:394-420 (damping=0)
        newton += 1
    if trace is not None:
        trace.setdefault("newton_counts", []).append(newton)
        trace.setdefault("newton_res", []).append(res_hist)
    return V


# --------------------------------------------------------------------------------------------
# uvsolve (src:433-589)
# --------------------------------------------------------------------------------------------
def uv_system(mesh, rec_mode, V, bddata, tau_p, tau_n, mu, n, u0, v0):
    """Operator A = M - J0 and right-hand side b of one uvsolve call (src:433-587), in the callee's
    own naming: G calls it as (V, ubddata, tau_p, tau_n, mu_p, u0, v0) for u and as
    (-V, vbddata, tau_n, tau_p, mu_n, v0, u0) for v (src:598-599)."""
    nq = mesh.nodes.shape[1]
    a1, a2, a3, uu, vv, ww, ar = _geometry(mesh)                                # src:443-456
    n0 = np.asarray(n) - 1
    X, Y = mesh.nodes
    V1, V2, V3 = V[a1], V[a2], V[a3]
    cx = (X[a1] + X[a2] + X[a3]) / 3                                            # src:463-464
    cy = (Y[a1] + Y[a2] + Y[a3]) / 3
    ph = _vec2(mu, cx, cy) * np.exp((V1 + V2 + V3) / 3)                         # src:467-469
    Kg = _stiffness_triplets(uu, vv, ww, ar, ph, ph).T.ravel()                  # src:471-483
    Ig, Jg = _local_ij(a1, a2, a3)
    u01, u02, u03 = u0[a1], u0[a2], u0[a3]
    v01, v02, v03 = v0[a1], v0[a2], v0[a3]
    Va, Vb, Vc = (V1 + V2) / 2, (V2 + V3) / 2, (V3 + V1) / 2                    # src:504-506
    v0a, v0b, v0c = (v01 + v02) / 2, (v02 + v03) / 2, (v03 + v01) / 2
    u0a, u0b, u0c = (u01 + u02) / 2, (u02 + u03) / 2, (u03 + u01) / 2
    if rec_mode == "shockley":                                                  # src:516-519
        fa = 1.0 / (tau_p * (np.exp(Va) * u0a + 1) + tau_n * (np.exp(-Va) * v0a + 1))
        fb = 1.0 / (tau_p * (np.exp(Vb) * u0b + 1) + tau_n * (np.exp(-Vb) * v0b + 1))
        fc = 1.0 / (tau_p * (np.exp(Vc) * u0c + 1) + tau_n * (np.exp(-Vc) * v0c + 1))
    elif rec_mode == "zero":                                                    # src:520-523
        fa = fb = fc = np.zeros(ar.size)
    else:
        raise ValueError("rec_mode must be 'shockley' or 'zero'")
    b = _load_vector(a1, a2, a3, fa, fb, fc, ar, nq)                            # src:525-530,562
    _, gD = dirichlet_data(mesh, bddata)                                        # src:532-544 (values from THIS bddata)
    if len(gD) != len(n0):
        raise ValueError("Dirichlet node set of u/v bddata differs from the V bddata (reference: DimensionMismatch at src:563)")
    M, bdM = _eliminate(Ig, Jg, Kg, n0, nq)                                     # src:546-559
    b = b - bdM[:, n0] @ gD                                                     # src:563
    b[n0] = gD                                                                  # src:564
    if rec_mode == "shockley":                                                  # src:566-584
        W1 = -v01 / (tau_p * (np.exp(V1) * u01 + 1) + tau_n * (np.exp(-V1) * v01 + 1)) * ar / 30
        W2 = -v02 / (tau_p * (np.exp(V2) * u02 + 1) + tau_n * (np.exp(-V2) * v02 + 1)) * ar / 30
        W3 = -v03 / (tau_p * (np.exp(V3) * u03 + 1) + tau_n * (np.exp(-V3) * v03 + 1)) * ar / 30
        J0 = _sparse(Ig, Jg, _wmass_triplets(W1, W2, W3).T.ravel(), nq)
        A = M - J0
    else:                                                                       # src:585-587
        A = M
    return A, b


def uvsolve(mesh, rec_mode, V, bddata, tau_p, tau_n, mu, n, u0, v0, linsolve="sparse", trace=None, tag=""):
    A, b = uv_system(mesh, rec_mode, V, bddata, tau_p, tau_n, mu, n, u0, v0)
    if trace is not None:
        trace["A_last" + tag] = A
        trace["b_last" + tag] = b
    return _solve(A, b, linsolve)                                               # src:588


# --------------------------------------------------------------------------------------------
# G and solve_ddpe (src:596-601, src:630-658)
# --------------------------------------------------------------------------------------------
def G(mesh, rec_mode, u0, v0, delta, tau_p, tau_n, mu_p, mu_n, gD, n, MV, bdMV, ar, ubddata, vbddata, C, tol,
      linsolve="sparse", trace=None, damping=0):
    V = Vsolve(mesh, u0, v0, delta, MV, bdMV, ar, n, gD, C, tol, linsolve, trace, damping=damping)     # src:597
    u = uvsolve(mesh, rec_mode, V, ubddata, tau_p, tau_n, mu_p, n, u0, v0, linsolve, trace, "_u")      # src:598
    v = uvsolve(mesh, rec_mode, -V, vbddata, tau_n, tau_p, mu_n, n, v0, u0, linsolve, trace, "_v")     # src:599 (old u0)
    return u, v, V


def solve_ddpe(mesh, rec_mode, Vbddata, ubddata, vbddata, lam, delta, tau_p, tau_n, mu_p, mu_n, C, tol=1e-9,
               linsolve="sparse", trace=None, max_gummel=200, verbose=False, damping=0):
    nq = mesh.nodes.shape[1]
    u = np.zeros(nq)
    v = np.zeros(nq)
    V = np.zeros(nq)                                                            # src:635-637
    _, _, gD, n = aquire_boundary(mesh, Vbddata)                                # src:640
    MV, bdMV, ar = MV_assemble(mesh, n, gD, lam)[:3]                            # src:643
    if trace is not None:
        trace["MV"] = MV
        trace["n"] = n
        trace["gD_V"] = gD
        trace["control"] = []
    control, count = 1.0, 0
    while control > tol and count < max_gummel:                                 # src:649
        u0, v0, V0 = u, v, V
        u, v, V = G(mesh, rec_mode, u0, v0, delta, tau_p, tau_n, mu_p, mu_n, gD, n, MV, bdMV, ar,
                    ubddata, vbddata, C, tol, linsolve, trace, damping)         # src:651
        control = max(np.max(np.abs(u - u0)), np.max(np.abs(v - v0)), np.max(np.abs(V - V0)))   # src:653
        count += 1
        if trace is not None:
            trace["control"].append(control)
        if verbose:
            print(f"iteration {count}, tolerance: {control}")                  # src:655
    return V, u, v


# --------------------------------------------------------------------------------------------
# solve_semlin_poisson (src:895-1097)
# --------------------------------------------------------------------------------------------
def solve_semlin_poisson(mesh, A, bddata, f, g, tol=1e-14, linsolve="sparse", trace=None, max_newton=200,
                         verbose=False, damping=0):
    nq = mesh.nodes.shape[1]
    gtag = mesh.elements[4, :].astype(np.float64)
    ph11 = np.array([float(A[0](t)) for t in gtag])                             # src:901
    ph22 = np.array([float(A[3](t)) for t in gtag])                             # src:902
    a1, a2, a3, uu, vv, ww, ar = _geometry(mesh)                                # src:910-921
    Kg = _stiffness_triplets(uu, vv, ww, ar, ph22, ph11).T.ravel()              # src:926 [ph22 ph11]
    Ig, Jg = _local_ij(a1, a2, a3)
    X, Y = mesh.nodes
    hax, hay = (X[a1] + X[a2]) / 2, (Y[a1] + Y[a2]) / 2                         # src:940-945
    hbx, hby = (X[a2] + X[a3]) / 2, (Y[a2] + Y[a3]) / 2
    hcx, hcy = (X[a3] + X[a1]) / 2, (Y[a3] + Y[a1]) / 2
    n1, n2, neuload = neumann_data(mesh, bddata)                                # src:976-996
    ex_idx = np.concatenate([n1 - 1, n2 - 1])
    ex_val = np.concatenate([neuload, neuload])
    n, gD = dirichlet_data(mesh, bddata)                                        # src:1001-1013
    n0 = n - 1
    M, bdM = _eliminate(Ig, Jg, Kg, n0, nq)                                     # src:1016-1028
    lift = bdM[:, n0] @ gD

    def rhs(V):
        V1, V2, V3 = V[a1], V[a2], V[a3]
        Ca, Cb, Cc = (V1 + V2) / 2, (V2 + V3) / 2, (V3 + V1) / 2                # src:1070-1072
        fa, fb, fc = _vec3(f, hax, hay, Ca), _vec3(f, hbx, hby, Cb), _vec3(f, hcx, hcy, Cc)
        b = _load_vector(a1, a2, a3, fa, fb, fc, ar, nq, ex_idx, ex_val)        # src:1077-1087
        b = b - lift                                                            # src:1088
        b[n0] = gD                                                              # src:1089
        return b

    V = np.zeros(nq)                                                            # src:948
    b = rhs(V)                                                                  # src:955-963,1031-1033
    if trace is not None:
        trace["M"] = M
        trace["n"] = n
        trace["gD"] = gD
        trace["b0"] = b.copy()
        trace["res"] = []
    count = 0
    while True:
        res = np.max(np.abs(M @ V - b))                                         # src:1040
        if trace is not None:
            trace["res"].append(res)
        if not res > tol or count >= max_newton:
            break
        W1 = _vec3(g, X[a1], Y[a1], V[a1]) * ar / 30                            # src:1046-1048 (signed ar)
        W2 = _vec3(g, X[a2], Y[a2], V[a2]) * ar / 30
        W3 = _vec3(g, X[a3], Y[a3], V[a3]) * ar / 30
        J0 = _sparse(Ig, Jg, _wmass_triplets(W1, W2, W3).T.ravel(), nq)         # src:1050-1062
        Aop = M - J0
        if trace is not None and "A_first" not in trace:
            trace["A_first"] = Aop.copy()
        V, b = _damped_update(V, _solve(Aop, M @ V - b, linsolve), res, M, rhs, damping)   # src:1063-1089 (damping=0)
        count += 1
        if verbose:
            print(f"iteration {count}, tolerance: {np.max(np.abs(M @ V - b))}")  # src:1091-1093
    if trace is not None:
        trace["newton"] = count
        trace["b_last"] = b
    return V


def poisson_newton_system(mesh, V, u, v, delta, M, bdM, ar, n, gD, C):
    """At the iterate V of Vsolve (src:368-420): Jacobian A = M - J0(V), rhs b(V), residual F = M V - b."""
    nq = mesh.nodes.shape[1]
    a1, a2, a3 = (mesh.elements[k, :] - 1 for k in range(3))
    n0 = np.asarray(n) - 1
    X, Y = mesh.nodes
    hax, hay = (X[a1] + X[a2]) / 2, (Y[a1] + Y[a2]) / 2
    hbx, hby = (X[a2] + X[a3]) / 2, (Y[a2] + Y[a3]) / 2
    hcx, hcy = (X[a3] + X[a1]) / 2, (Y[a3] + Y[a1]) / 2
    u1, u2, u3 = u[a1], u[a2], u[a3]
    v1, v2, v3 = v[a1], v[a2], v[a3]
    ua, ub, uc = (u1 + u2) / 2, (u2 + u3) / 2, (u3 + u1) / 2
    va, vb, vc = (v1 + v2) / 2, (v2 + v3) / 2, (v3 + v1) / 2
    V1, V2, V3 = V[a1], V[a2], V[a3]
    Sa, Sb, Sc = (V1 + V2) / 2, (V2 + V3) / 2, (V3 + V1) / 2
    fa = delta ** 2 * (np.exp(-Sa) * va - np.exp(Sa) * ua) + _vec2(C, hax, hay)
    fb = delta ** 2 * (np.exp(-Sb) * vb - np.exp(Sb) * ub) + _vec2(C, hbx, hby)
    fc = delta ** 2 * (np.exp(-Sc) * vc - np.exp(Sc) * uc) + _vec2(C, hcx, hcy)
    b = _load_vector(a1, a2, a3, fa, fb, fc, ar, nq) - bdM[:, n0] @ gD
    b[n0] = gD
    W1 = -delta ** 2 * (u1 * np.exp(V1) + v1 * np.exp(-V1)) * ar / 30
    W2 = -delta ** 2 * (u2 * np.exp(V2) + v2 * np.exp(-V2)) * ar / 30
    W3 = -delta ** 2 * (u3 * np.exp(V3) + v3 * np.exp(-V3)) * ar / 30
    Ig, Jg = _local_ij(a1, a2, a3)
    J0 = _sparse(Ig, Jg, _wmass_triplets(W1, W2, W3).T.ravel(), nq)
    return M - J0, b, M @ V - b


def semlin_newton_system(mesh, A, bddata, f, g, V):
    """At the iterate V of solve_semlin_poisson (src:1040-1089): (M, A = M - J0(V), b(V), F)."""
    nq = mesh.nodes.shape[1]
    gtag = mesh.elements[4, :].astype(np.float64)
    ph11 = np.array([float(A[0](t)) for t in gtag])
    ph22 = np.array([float(A[3](t)) for t in gtag])
    a1, a2, a3, uu, vv, ww, ar = _geometry(mesh)
    Kg = _stiffness_triplets(uu, vv, ww, ar, ph22, ph11).T.ravel()
    Ig, Jg = _local_ij(a1, a2, a3)
    X, Y = mesh.nodes
    n1, n2, neuload = neumann_data(mesh, bddata)
    n, gD = dirichlet_data(mesh, bddata)
    n0 = n - 1
    M, bdM = _eliminate(Ig, Jg, Kg, n0, nq)
    V1, V2, V3 = V[a1], V[a2], V[a3]
    fa = _vec3(f, (X[a1] + X[a2]) / 2, (Y[a1] + Y[a2]) / 2, (V1 + V2) / 2)
    fb = _vec3(f, (X[a2] + X[a3]) / 2, (Y[a2] + Y[a3]) / 2, (V2 + V3) / 2)
    fc = _vec3(f, (X[a3] + X[a1]) / 2, (Y[a3] + Y[a1]) / 2, (V3 + V1) / 2)
    b = _load_vector(a1, a2, a3, fa, fb, fc, ar, nq, np.concatenate([n1 - 1, n2 - 1]), np.concatenate([neuload, neuload]))
    b = b - bdM[:, n0] @ gD
    b[n0] = gD
    W1 = _vec3(g, X[a1], Y[a1], V1) * ar / 30
    W2 = _vec3(g, X[a2], Y[a2], V2) * ar / 30
    W3 = _vec3(g, X[a3], Y[a3], V3) * ar / 30
    J0 = _sparse(Ig, Jg, _wmass_triplets(W1, W2, W3).T.ravel(), nq)
    return M, M - J0, b, M @ V - b


# --------------------------------------------------------------------------------------------
# calculate_current (src:662-887) -- literal restatement incl. the O(ne*nme) findfirst
# --------------------------------------------------------------------------------------------
def _jl_isapprox(a, b):
    return abs(a - b) <= np.sqrt(np.finfo(float).eps) * max(abs(a), abs(b))


def aquire_boundary_separate_edges(mesh, Endpoints, bddata):
    """src:662-727."""
    E = np.asarray(Endpoints, dtype=float)
    n1, n2, gD, n = aquire_boundary(mesh, bddata)
    tags, fcts, _, diri = _bd_columns(bddata)
    idx = []
    for j in diri:
        idx.extend(np.nonzero(mesh.edges[3, :] == tags[j])[0].tolist())          # src:690-691
    X, Y = mesh.nodes
    ed = mesh.edges
    xpar = [k for k, e in enumerate(idx) if _jl_isapprox(Y[ed[0, e] - 1] - Y[ed[1, e] - 1], 0.0)]   # src:701
    ypar = [k for k, e in enumerate(idx) if _jl_isapprox(X[ed[0, e] - 1] - X[ed[1, e] - 1], 0.0)]   # src:702
    assert len(xpar) + len(ypar) == len(idx)
    left = [k for k in ypar if _jl_isapprox(E[0, 0], X[ed[0, idx[k]] - 1])]                         # src:708
    bottom = [k for k in xpar if _jl_isapprox(E[0, 1], Y[ed[0, idx[k]] - 1])]
    right = [k for k in ypar if _jl_isapprox(E[2, 0], X[ed[0, idx[k]] - 1])]
    top = [k for k in xpar if _jl_isapprox(E[2, 1], Y[ed[0, idx[k]] - 1])]
    assert len(left) + len(bottom) + len(right) + len(top) == len(idx)
    pick = lambda ks: ed[:, [idx[k] for k in ks]]
    return n1, n2, gD, n, pick(left), pick(bottom), pick(right), pick(top)


def calculate_current(mesh, rec_mode, Endpoints, Vbddata, ubddata, vbddata, lam, delta, tau_p, tau_n, mu_p, mu_n, C,
                      tol=1e-9, linsolve="sparse", fields=None):
    """src:753-887.  ``fields=(V,u,v)`` skips the Gummel loop (used to test the flux integrals alone)."""
    E = np.asarray(Endpoints, dtype=float)
    _, _, gD, n, eL, eB, eR, eT = aquire_boundary_separate_edges(mesh, E, Vbddata)
    if fields is None:
        V, u, v = solve_ddpe(mesh, rec_mode, Vbddata, ubddata, vbddata, lam, delta, tau_p, tau_n, mu_p, mu_n, C, tol, linsolve)
    else:
        V, u, v = fields
    a1, a2, a3, uu, vv, ww, ar = _geometry(mesh)
    gradx = np.stack([.5 * uu[1] / ar, .5 * vv[1] / ar, .5 * ww[1] / ar], axis=1)        # src:737-740
    grady = np.stack([-.5 * uu[0] / ar, -.5 * vv[0] / ar, -.5 * ww[0] / ar], axis=1)     # src:743-746
    gl_nodes = np.array([-np.sqrt(3 / 5), 0.0, np.sqrt(3 / 5)])                          # gausslegendre(3), src:792
    gl_w = np.array([5 / 9, 8 / 9, 5 / 9])
    tri = mesh.elements[0:3, :]
    I = 0.0
    for edges, fixed, along_x, gtab, sign in ((eL, E[0, 0], False, gradx, -1.0), (eB, E[0, 1], True, grady, -1.0),
                                              (eR, E[2, 0], False, gradx, 1.0), (eT, E[2, 1], True, grady, 1.0)):
        for i in range(edges.shape[1]):
            Vb = Vbddata[2][int(edges[3, i]) - 1]                                        # src:801
            c = mesh.nodes[0 if along_x else 1, edges[0:2, i] - 1]
            a, b = c.min(), c.max()
            ind = next(x for x in range(tri.shape[1]) if set(edges[0:2, i]) <= set(tri[:, x]))   # src:807
            gradu = np.dot(u[tri[:, ind] - 1], gtab[ind])
            gradv = np.dot(v[tri[:, ind] - 1], gtab[ind])
            Ip = np.zeros(3)
            In = np.zeros(3)
            for q, t in enumerate(gl_nodes):
                s = (t + 1) * (b - a) / 2 + a                                            # src:798
                x, y = (s, fixed) if along_x else (fixed, s)
                Ip[q] = float(mu_n(x, y)) * np.exp(float(Vb(x, y))) * gradu * (b - a) / 2      # src:802,814
                In[q] = -float(mu_p(x, y)) * np.exp(-float(Vb(x, y))) * gradv * (b - a) / 2   # src:803,815
            I += sign * np.dot(gl_w, Ip + In)                                            # src:817,839,861,883
    return I, V, u, v


# --------------------------------------------------------------------------------------------
# Timing helper for bench.py's cpu_baseline leg
# --------------------------------------------------------------------------------------------
def timed_gummel_iterations(mesh, problem: dict, iters: int, linsolve="sparse"):
    """Run ``iters`` Gummel iterations (src:649-656) from the zero state and return seconds per
    iteration.  ``problem`` holds the solve_ddpe arguments."""
    nq = mesh.nodes.shape[1]
    u = np.zeros(nq)
    v = np.zeros(nq)
    _, _, gD, n = aquire_boundary(mesh, problem["Vbddata"])
    MV, bdMV, ar = MV_assemble(mesh, n, gD, problem["lam"])[:3]
    t0 = time.perf_counter()
    for _ in range(iters):
        u, v, V = G(mesh, problem["rec_mode"], u, v, problem["delta"], problem["tau_p"], problem["tau_n"],
                    problem["mu_p"], problem["mu_n"], gD, n, MV, bdMV, ar, problem["ubddata"],
                    problem["vbddata"], problem["C"], problem["tol"], linsolve)
    return (time.perf_counter() - t0) / iters
