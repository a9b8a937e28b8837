#!/usr/bin/env python
"""CPU prototype (numpy/scipy) used to CHOOSE the multigrid preconditioner design of SURVEY 8f rank 1
before writing CUDA: plain-aggregation AMG (piecewise-constant prolongation, Galerkin coarse
operators = sums of fine entries) as a preconditioner of CG, on the operators of the bench problem
(pn-diode on a structured triangulation, built by the oracle).  Test infrastructure, not product.

  python tools/amg_prototype.py 512 256
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ddps_oracle as O  # noqa: E402
import ddps_b200 as D  # noqa: E402


def hash32(i, seed):
    x = (i.astype(np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B9)) & np.uint64(0xFFFFFFFF)
    x = ((x ^ (x >> np.uint64(16))) * np.uint64(0x45D9F3B)) & np.uint64(0xFFFFFFFF)
    x = ((x ^ (x >> np.uint64(16))) * np.uint64(0x45D9F3B)) & np.uint64(0xFFFFFFFF)
    return (x ^ (x >> np.uint64(16))).astype(np.int64)


def pairwise_pass(A, active, rounds=4, seed=0):
    """Handshake matching on the strength graph of A.  Returns agg (n) with -1 for inactive nodes."""
    n = A.shape[0]
    A = A.tocsr()
    indptr, indices, data = A.indptr, A.indices, A.data
    rows = np.repeat(np.arange(n), np.diff(indptr))
    off = (rows != indices) & active[rows] & active[indices]
    rows, cols, w = rows[off], indices[off], -data[off]          # M-matrix: strong = large negative
    keep = w > 0
    rows, cols, w = rows[keep], cols[keep], w[keep]
    # relative strength threshold
    mx = np.zeros(n)
    np.maximum.at(mx, rows, w)
    strong = w >= 0.25 * mx[rows]
    rows, cols, w = rows[strong], cols[strong], w[strong]
    mate = -np.ones(n, dtype=np.int64)
    for r in range(rounds):
        free = (mate[rows] < 0) & (mate[cols] < 0)
        rr, cc, ww = rows[free], cols[free], w[free]
        if rr.size == 0:
            break
        # pick strongest free neighbour, ties by hash
        key = hash32(np.minimum(rr, cc) * 7919 + np.maximum(rr, cc), seed + r).astype(np.float64)
        order = np.lexsort((cc, -key, rr))
        rr_s, cc_s = rr[order], cc[order]
        first = np.ones(rr_s.size, dtype=bool)
        first[1:] = rr_s[1:] != rr_s[:-1]
        pick = -np.ones(n, dtype=np.int64)
        pick[rr_s[first]] = cc_s[first]
        i = np.nonzero(pick >= 0)[0]
        mutual = i[pick[pick[i]] == i]
        mate[mutual] = pick[mutual]
    # aggregates: pairs get the id of min(i,mate); leftovers join strongest neighbour's pair (if any) else singleton
    rep = np.where(mate >= 0, np.minimum(np.arange(n), mate), -1)
    left = np.nonzero((mate < 0) & active)[0]
    if left.size:
        isleft = np.zeros(n, dtype=bool)
        isleft[left] = True
        sel = isleft[rows] & (mate[cols] >= 0)
        rr, cc, ww = rows[sel], cols[sel], w[sel]
        order = np.lexsort((cc, -ww, rr))
        rr_s, cc_s = rr[order], cc[order]
        first = np.ones(rr_s.size, dtype=bool)
        first[1:] = rr_s[1:] != rr_s[:-1]
        rep[rr_s[first]] = rep[cc_s[first]]
        still = left[rep[left] < 0]
        rep[still] = still
    ids = np.unique(rep[rep >= 0])
    remap = -np.ones(n, dtype=np.int64)
    remap[ids] = np.arange(ids.size)
    agg = np.where(rep >= 0, remap[np.maximum(rep, 0)], -1)
    return agg, ids.size


def mis2_aggregate(A, active, seed=0):
    """Root = distance-2 independent set (Luby with hashed priorities); aggregate = root + its neighbours;
    leftovers join the neighbouring aggregate of the strongest connection."""
    n = A.shape[0]
    A = A.tocsr()
    indptr, indices, data = A.indptr, A.indices, A.data
    rows = np.repeat(np.arange(n), np.diff(indptr))
    off = (rows != indices) & active[rows] & active[indices] & (data < 0)
    rows, cols, w = rows[off], indices[off], -data[off]
    mx = np.zeros(n)
    np.maximum.at(mx, rows, w)
    strong = w >= 0.25 * mx[rows]
    rows, cols, w = rows[strong], cols[strong], w[strong]
    prio = hash32(np.arange(n), seed) * n + np.arange(n)
    state = np.where(active, 0, 2)  # 0 undecided, 1 root, 2 excluded
    for it in range(50):
        und = state == 0
        if not und.any():
            break
        p = np.where(und, prio, -1)
        p[state == 1] = np.iinfo(np.int64).max
        m1 = p.copy()
        np.maximum.at(m1, rows, p[cols])
        m2 = m1.copy()
        np.maximum.at(m2, rows, m1[cols])
        newroot = und & (m2 == prio)
        state[newroot] = 1
        # exclude nodes within distance 2 of a root
        r1 = np.zeros(n, dtype=bool)
        r1[rows[state[cols] == 1]] = True
        r2 = r1.copy()
        r2[rows[r1[cols]]] = True
        state[(state == 0) & r2] = 2
    roots = np.nonzero(state == 1)[0]
    agg = -np.ones(n, dtype=np.int64)
    agg[roots] = np.arange(roots.size)
    sel = (state[cols] == 1) & (agg[rows] < 0)
    # neighbours of roots (each node has at most one root neighbour? no: distance-2 MIS => unique)
    agg[rows[sel]] = agg[cols[sel]]
    left = np.nonzero((agg < 0) & active)[0]
    if left.size:
        isleft = np.zeros(n, dtype=bool)
        isleft[left] = True
        base = agg.copy()
        sel = isleft[rows] & (base[cols] >= 0)
        rr, cc, ww = rows[sel], cols[sel], w[sel]
        order = np.lexsort((cc, -ww, rr))
        rr_s, cc_s = rr[order], cc[order]
        first = np.ones(rr_s.size, dtype=bool)
        first[1:] = rr_s[1:] != rr_s[:-1]
        agg[rr_s[first]] = base[cc_s[first]]
        still = left[agg[left] < 0]
        agg[still] = roots.size + np.arange(still.size)
    nc = int(agg.max()) + 1
    return agg, nc


def build_hierarchy(A, isD, kind="pair2", min_coarse=200, max_levels=20):
    levels = []
    active = ~isD
    Al = A.tocsr()
    while True:
        n = Al.shape[0]
        dinv = 1.0 / Al.diagonal()
        lev = dict(A=Al, dinv=dinv)
        levels.append(lev)
        if n <= min_coarse or len(levels) >= max_levels:
            break
        if kind == "pair2":
            a1, n1 = pairwise_pass(Al, active, seed=len(levels))
            T1 = tent(a1, n1)
            A1 = (T1.T @ Al @ T1).tocsr()
            a2, n2 = pairwise_pass(A1, np.ones(n1, dtype=bool), seed=100 + len(levels))
            agg = np.where(a1 >= 0, a2[np.maximum(a1, 0)], -1)
            nc = n2
        elif kind == "pair3":
            a1, n1 = pairwise_pass(Al, active, seed=len(levels))
            T1 = tent(a1, n1)
            A1 = (T1.T @ Al @ T1).tocsr()
            a2, n2 = pairwise_pass(A1, np.ones(n1, dtype=bool), seed=100 + len(levels))
            T2 = tent(a2, n2)
            A2 = (T2.T @ A1 @ T2).tocsr()
            a3, n3 = pairwise_pass(A2, np.ones(n2, dtype=bool), seed=200 + len(levels))
            a12 = np.where(a1 >= 0, a2[np.maximum(a1, 0)], -1)
            agg = np.where(a12 >= 0, a3[np.maximum(a12, 0)], -1)
            nc = n3
        else:
            agg, nc = mis2_aggregate(Al, active, seed=len(levels))
        if nc >= 0.8 * n:
            break
        T = tent(agg, nc)
        lev["T"] = T
        lev["agg"] = agg
        Al = (T.T @ Al @ T).tocsr()
        active = np.ones(nc, dtype=bool)
    return levels


def tent(agg, nc):
    i = np.nonzero(agg >= 0)[0]
    return sp.csr_matrix((np.ones(i.size), (i, agg[i])), shape=(agg.size, nc))


def vcycle(levels, l, b, nu=1, omega=0.7, gamma=1.0, coarse_sweeps=30, cheb=None):
    lev = levels[l]
    A, dinv = lev["A"], lev["dinv"]
    if "T" not in lev:
        if A.shape[0] <= 1500:
            if "dense" not in lev:
                lev["dense"] = np.linalg.inv(A.toarray())
            return lev["dense"] @ b
        x = omega * dinv * b
        for _ in range(coarse_sweeps - 1):
            x += omega * dinv * (b - A @ x)
        return x
    x = omega * dinv * b
    for _ in range(nu - 1):
        x += omega * dinv * (b - A @ x)
    r = b - A @ x
    bc = lev["T"].T @ r
    ec = vcycle(levels, l + 1, bc, nu, omega, gamma, coarse_sweeps)
    g = gamma if np.isscalar(gamma) else gamma[min(l, len(gamma) - 1)]
    x += g * (lev["T"] @ ec)
    for _ in range(nu):
        x += omega * dinv * (b - A @ x)
    return x


def kcycle(levels, l, b, nu=1, omega=0.7, kdepth=99):
    """K-cycle: 2 FCG iterations at each coarse level (Notay)."""
    lev = levels[l]
    A, dinv = lev["A"], lev["dinv"]
    if "T" not in lev:
        if "dense" not in lev:
            lev["dense"] = np.linalg.inv(A.toarray())
        return lev["dense"] @ b
    x = omega * dinv * b
    for _ in range(nu - 1):
        x += omega * dinv * (b - A @ x)
    r = b - A @ x
    bc = lev["T"].T @ r
    Ac = levels[l + 1]["A"]
    if l + 1 >= kdepth or "T" not in levels[l + 1]:
        ec = kcycle(levels, l + 1, bc, nu, omega, kdepth)
    else:
        c = kcycle(levels, l + 1, bc, nu, omega, kdepth)
        v = Ac @ c
        rho1 = c @ v
        a1 = c @ bc
        rc = bc - (a1 / rho1) * v
        if np.linalg.norm(rc) <= 0.25 * np.linalg.norm(bc):
            ec = (a1 / rho1) * c
        else:
            d = kcycle(levels, l + 1, rc, nu, omega, kdepth)
            w = Ac @ d
            g = d @ v
            be = d @ w
            a2 = d @ rc
            rho2 = be - g * g / rho1
            ec = (a1 / rho1 - g * a2 / (rho1 * rho2)) * c + (a2 / rho2) * d
    x += lev["T"] @ ec
    for _ in range(nu):
        x += omega * dinv * (b - A @ x)
    return x


def pcg(A, b, M, rtol=1e-13, maxit=5000, flexible=False):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = r @ z
    bb = np.sqrt(b @ b)
    for it in range(1, maxit + 1):
        q = A @ p
        al = rz / (p @ q)
        x += al * p
        rold = r.copy()
        r -= al * q
        if np.sqrt(r @ r) <= rtol * bb:
            return x, it
        z = M(r)
        if flexible:
            rz_new = z @ (r - rold)
            be = rz_new / rz
            rz = r @ z
        else:
            rz_new = r @ z
            be = rz_new / rz
            rz = rz_new
        p = z + be * p
    return x, maxit


def operators(nx, ny):
    mesh = O.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0)
    pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
    nq = mesh.nodes.shape[1]
    _, _, gD, n = O.aquire_boundary(mesh, pr["Vbddata"])
    MV, bdMV, ar = O.MV_assemble(mesh, n, gD, pr["lam"])[:3]
    # a plausible state: one Gummel-ish field from a smooth V profile
    X = mesh.nodes[0]
    V = 0.6 * np.tanh(4 * (X - 1.0)) + 0.15 * (X > 1)
    u0 = np.exp(-0.3 * (X > 1.0) * (X - 1))
    v0 = np.exp(0.3 * (X > 1.0) * (X - 1))
    A, b = O.uv_system(mesh, "shockley", V, pr["ubddata"], 1.0, 1.0, pr["mu_p"], n, u0, v0)
    isD = np.zeros(nq, dtype=bool)
    isD[np.asarray(n) - 1] = True
    return mesh, MV.tocsr(), A.tocsr(), b, isD


def main():
    nx, ny = int(sys.argv[1]), int(sys.argv[2])
    t0 = time.time()
    mesh, MV, A, b, isD = operators(nx, ny)
    print(f"mesh {nx}x{ny}: nq={A.shape[0]} nnz={A.nnz} build {time.time()-t0:.1f}s", flush=True)
    dinv = 1.0 / A.diagonal()
    t0 = time.time()
    x, it = pcg(A, b, lambda r: dinv * r, maxit=20000)
    print(f"jacobi-PCG: {it} its  ({time.time()-t0:.1f}s)", flush=True)
    for kind in sys.argv[3].split(",") if len(sys.argv) > 3 else ["pair2", "mis2", "pair3"]:
        t0 = time.time()
        levels = build_hierarchy(A, isD, kind)
        sizes = [l["A"].shape[0] for l in levels]
        nnzs = [l["A"].nnz for l in levels]
        print(f"[{kind}] levels {sizes}\n    nnz {nnzs} opcx={sum(nnzs)/nnzs[0]:.2f}  setup {time.time()-t0:.1f}s", flush=True)
        for nu, om, ga in [(1, 0.7, 1.0), (1, 0.7, 1.5), (1, 0.7, 1.8), (2, 0.7, 1.5), (2, 0.7, 1.8), (1, 0.8, 1.8)]:
            t0 = time.time()
            x2, it2 = pcg(A, b, lambda r: vcycle(levels, 0, r, nu, om, ga), maxit=600)
            err = np.linalg.norm(x2 - x) / np.linalg.norm(x)
            print(f"   V nu={nu} om={om} gamma={ga}: {it2} its  err={err:.1e} ({time.time()-t0:.1f}s)", flush=True)
        for nu in (1, 2):
            t0 = time.time()
            x2, it2 = pcg(A, b, lambda r: kcycle(levels, 0, r, nu, 0.7), maxit=300, flexible=True)
            err = np.linalg.norm(x2 - x) / np.linalg.norm(x)
            print(f"   K nu={nu}: {it2} its err={err:.1e} ({time.time()-t0:.1f}s)", flush=True)




def build_sa(A, isD, L=None, omega=2.0 / 3.0, min_coarse=200, kind="mis2", filt=False):
    """Smoothed aggregation; if L is given the prolongators are built from L (frozen) instead of A."""
    levels = []
    active = ~isD
    Al = A.tocsr()
    Ll = (L if L is not None else A).tocsr()
    while True:
        n = Al.shape[0]
        lev = dict(A=Al, dinv=1.0 / Al.diagonal())
        levels.append(lev)
        if n <= min_coarse:
            break
        if kind == "mis2":
            agg, nc = mis2_aggregate(Ll, active, seed=len(levels))
        else:
            a1, n1 = pairwise_pass(Ll, active, seed=len(levels))
            T1 = tent(a1, n1)
            A1 = (T1.T @ Ll @ T1).tocsr()
            a2, n2 = pairwise_pass(A1, np.ones(n1, dtype=bool), seed=100 + len(levels))
            agg = np.where(a1 >= 0, a2[np.maximum(a1, 0)], -1)
            nc = n2
        if nc >= 0.8 * n:
            break
        T = tent(agg, nc)
        Dl = sp.diags(1.0 / Ll.diagonal())
        S = sp.identity(n) - omega * (Dl @ Ll)
        if active is not None and (~active).any():
            # Dirichlet rows: keep P row zero
            S = sp.diags(active.astype(float)) @ S
        P = (S @ T).tocsr()
        lev["T"] = P
        Al = (P.T @ Al @ P).tocsr()
        Ll = (P.T @ Ll @ P).tocsr() if L is not None else Al
        active = np.ones(nc, dtype=bool)
    return levels


def main_sa():
    nx, ny = int(sys.argv[2]), int(sys.argv[3])
    mesh, MV, A, b, isD = operators(nx, ny)
    print(f"mesh {nx}x{ny}: nq={A.shape[0]} nnz={A.nnz}", flush=True)
    dinv = 1.0 / A.diagonal()
    x, it = pcg(A, b, lambda r: dinv * r, maxit=20000)
    print(f"jacobi-PCG: {it} its", flush=True)
    for name, L, kind in [("SA(A) mis2", None, "mis2"), ("SA(frozen MV) mis2", MV, "mis2"), ("SA(A) pair2", None, "pair2")]:
        t0 = time.time()
        levels = build_sa(A, isD, L, kind=kind)
        sizes = [l["A"].shape[0] for l in levels]
        nnzs = [l["A"].nnz for l in levels]
        pn = [l["T"].nnz / l["T"].shape[0] for l in levels if "T" in l]
        trip = [int((np.diff(l["T"].indptr)[l["A"].tocoo().row] * np.diff(l["T"].indptr)[l["A"].tocoo().col]).sum()) for l in levels if "T" in l]
        print(f"[{name}] levels {sizes}\n   nnz {nnzs} opcx={sum(nnzs)/nnzs[0]:.2f} P nnz/row {np.round(pn,2)} triple-products {trip} setup {time.time()-t0:.1f}s", flush=True)
        for nu, om in [(1, 0.7), (1, 0.8), (2, 0.7)]:
            x2, it2 = pcg(A, b, lambda r: vcycle(levels, 0, r, nu, om, 1.0), maxit=600)
            err = np.linalg.norm(x2 - x) / np.linalg.norm(x)
            print(f"   V nu={nu} om={om}: {it2} its  err={err:.1e}", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "sa":
    main_sa()
elif __name__ == "__main__":
    main()
