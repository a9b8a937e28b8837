"""CPU tests of the multigrid hierarchy setup: the host-only entry points of the library
(ddps_amg_host_*, no GPU) against scipy, on the reference's two gmsh meshes and a structured mesh, with the
operators of the CPU oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

import amg_ref
from util import oracle_mesh


@pytest.fixture(scope="module")
def D():
    import ddps_b200
    return ddps_b200


@pytest.fixture(scope="module")
def O():
    import ddps_oracle
    return ddps_oracle


def _systems(D, O, mesh, pr):
    """Base stiffness MV, a continuity operator A = M - J0 (with right-hand side) and the Dirichlet flags."""
    nq = mesh.nodes.shape[1]
    _, _, gD, n = O.aquire_boundary(mesh, pr["Vbddata"])
    MV = O.MV_assemble(mesh, n, gD, pr["lam"])[0].tocsr()
    X, Y = mesh.nodes
    V = 0.6 * np.tanh(3 * X) + 0.1 * Y
    u0 = np.exp(-0.3 * X)
    v0 = np.exp(0.2 * X * Y)
    A, b = O.uv_system(mesh, "shockley", V, pr["ubddata"], 1.0, 1.0, pr["mu_p"], n, u0, v0)
    fixed = np.zeros(nq, dtype=bool)
    fixed[np.asarray(n) - 1] = True
    MV.sort_indices()
    return MV, A.tocsr(), b, fixed


def _mesh_and_problem(D, O, which):
    if which == "F1":
        return oracle_mesh("F1"), D.problems.pn_junction([2, 4], [1, 2, 3, 4], C0=1.0, U=0.3)
    if which == "F2":
        return oracle_mesh("F2"), D.problems.pn_junction([3, 7], [1, 2, 3, 4, 5, 6, 7, 8], C0=1.0, U=0.3)
    m = O.structured_mesh(96, 48, 0.0, 2.0, 0.0, 1.0)
    return m, D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)


@pytest.mark.parametrize("which", ["F1", "F2", "S"])
def test_hierarchy_against_scipy(D, O, which):
    mesh, pr = _mesh_and_problem(D, O, which)
    MV, A, b, fixed = _systems(D, O, mesh, pr)
    levels = D.amg.host_hierarchy(MV.indptr, MV.indices, MV.data, fixed, coarse_max=60)
    assert len(levels) >= 3
    mats = amg_ref.hierarchy_matrices(levels)
    omega = 2.0 / 3.0
    for l, (Al, P) in enumerate(mats[:-1]):
        lev = levels[l]
        agg = lev["agg"]
        n, nc = lev["n"], lev["nc"]
        assert nc < 0.7 * n and agg.max() == nc - 1 and np.all(np.bincount(agg[agg >= 0], minlength=nc) >= 1)
        if l == 0:
            assert np.all(agg[fixed] == -1)                      # Dirichlet rows are never aggregated
            assert np.all(np.diff(P.indptr)[fixed] == 0)         # and their prolongator rows are empty
        # P = (I - omega D^-1 A) T with the piecewise-constant T of the aggregates
        i = np.nonzero(agg >= 0)[0]
        T = sp.csr_matrix((np.ones(i.size), (i, agg[i])), shape=(n, nc))
        S = sp.identity(n) - omega * sp.diags(1.0 / Al.diagonal()) @ Al
        if l == 0:
            S = sp.diags((~fixed).astype(float)) @ S
        ref = (S @ T).tocsr()
        d = (ref - P).tocoo()
        assert (np.max(np.abs(d.data)) if d.nnz else 0.0) <= 1e-7        # ... stored in fp32 on the device:
        assert np.array_equal(P.data, P.data.astype(np.float32).astype(np.float64))   # rounded once, on the host
        # coarse base operator = Galerkin product, sorted columns, pattern contains the structural product
        G = (P.T @ Al @ P).tocsr()
        Ac = mats[l + 1][0]
        dd = (G - Ac).tocoo()
        assert (np.max(np.abs(dd.data)) if dd.nnz else 0.0) <= 1e-12 * np.max(np.abs(Ac.data))
        assert np.all(np.diff(levels[l + 1]["col"])[np.diff(np.repeat(np.arange(nc), np.diff(levels[l + 1]["rowptr"]))) == 0] > 0)
        Gs = (abs(P).T @ (abs(Al) + sp.identity(n)) @ abs(P)).tocsr()   # structural pattern incl. explicit zeros
        Gs.data[:] = 1.0
        Acs = Ac.copy()
        Acs.data[:] = 1.0
        assert (Gs - Gs.multiply(Acs)).nnz == 0 or abs(Gs - Gs.multiply(Acs)).max() == 0


@pytest.mark.parametrize("which", ["F1", "F2", "S"])
def test_vcycle_pcg_converges_fast(D, O, which):
    mesh, pr = _mesh_and_problem(D, O, which)
    MV, A, b, fixed = _systems(D, O, mesh, pr)
    levels = D.amg.host_hierarchy(MV.indptr, MV.indices, MV.data, fixed, coarse_max=60)
    ops, Ps = amg_ref.operator_levels(A, levels)
    x, it = amg_ref.pcg(A, b, lambda r: amg_ref.vcycle(ops, Ps, r))
    dinv = 1.0 / A.diagonal()
    xj, itj = amg_ref.pcg(A, b, lambda r: dinv * r, maxit=5000)
    assert np.linalg.norm(x - xj) <= 1e-10 * np.linalg.norm(xj)
    assert it <= 30 and it * 5 < itj, (it, itj)


def test_thread_count_does_not_change_the_hierarchy(D, O):
    mesh, pr = _mesh_and_problem(D, O, "S")
    MV, A, b, fixed = _systems(D, O, mesh, pr)
    a = D.amg.host_hierarchy(MV.indptr, MV.indices, MV.data, fixed, nthreads=1, coarse_max=60)
    c = D.amg.host_hierarchy(MV.indptr, MV.indices, MV.data, fixed, nthreads=5, coarse_max=60)
    assert len(a) == len(c)
    for la, lc in zip(a, c):
        for k in la:
            assert np.array_equal(la[k], lc[k]), k


def test_prolongator_preserves_constants_and_gentle_last_levels(D, O):
    """Interior rows of the un-eliminated part of a stiffness matrix have zero row sums, so P must interpolate the constant
    exactly there (partition of unity, up to the fp32 rounding of the stored entries); and the last, tiny levels are
    coarsened gently (aggregates of <= 3 rows, ratio ~3 instead of ~7) so that the coarsest operator stays small without
    costing CG iterations."""
    mesh, pr = _mesh_and_problem(D, O, "S")
    MV, A, b, fixed = _systems(D, O, mesh, pr)
    levels = D.amg.host_hierarchy(MV.indptr, MV.indices, MV.data, fixed, coarse_max=12)
    mats = amg_ref.hierarchy_matrices(levels)
    A0, P0 = mats[0]
    rowsum = np.abs(np.asarray(A0.sum(axis=1)).ravel())
    touches_fixed = np.asarray((abs(A0) @ fixed.astype(float))).ravel() > 0
    agg = levels[0]["agg"]
    nb_unagg = np.asarray((abs(A0) @ (agg < 0).astype(float))).ravel() > 0
    interior = (~fixed) & (~touches_fixed) & (~nb_unagg) & (rowsum < 1e-12)
    assert interior.sum() > 0.8 * (~fixed).sum()
    ones = np.asarray(P0.sum(axis=1)).ravel()
    assert np.max(np.abs(ones[interior] - 1.0)) <= 5e-7
    # gentle coarsening below 512 rows: aggregates of at most 3 rows
    small = [lev for lev in levels if lev["nc"] and lev["n"] <= 512]
    assert small, [lev["n"] for lev in levels]
    for lev in small:
        sizes = np.bincount(lev["agg"][lev["agg"] >= 0], minlength=lev["nc"])
        assert sizes.max() <= 3 + 2 and lev["nc"] >= lev["n"] // 4      # root + 2, plus leftovers that joined
    big = [lev for lev in levels if lev["nc"] and lev["n"] > 512]
    assert all(lev["nc"] <= lev["n"] // 4 for lev in big)
