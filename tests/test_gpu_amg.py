"""GPU tests of the multigrid-preconditioned CG (opts.precond = 1), through the C ABI:
device Galerkin products / V-cycle against the numpy restatement (tests/amg_ref.py), solver parity with the CPU
oracle's golden fixtures (same tolerances and iteration counts as the Jacobi path), determinism."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import amg_ref
from util import KEYS, four_tag, golden, gpu_mesh, oracle_mesh, rel_l2, rel_max, smooth_fields

pytestmark = pytest.mark.gpu

FIELD_TOL = 1e-8


@pytest.fixture(scope="module")
def D():
    import ddps_b200
    return ddps_b200


@pytest.fixture(scope="module")
def O():
    import ddps_oracle
    return ddps_oracle


def _session(D, mesh, pr, **opts):
    s = D.Session(0, precond=1, **opts)
    s.set_mesh(mesh)
    s.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
    return s


def _fixed(D, mesh, pr):
    n, _ = D.dirichlet_nodes(mesh, pr["Vbddata"])
    f = np.zeros(mesh.nq, dtype=bool)
    f[np.asarray(n) - 1] = True
    return f


MESHES = {
    "F2": lambda D: (gpu_mesh("F2"), D.problems.pn_junction([3, 7], [1, 2, 3, 4, 5, 6, 7, 8], C0=1.0, U=0.3)),
    "S": lambda D: (D.structured_mesh(160, 80, 0.0, 2.0, 0.0, 1.0),
                    D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)),
}


def _same_hierarchy(dev, host):
    assert [lev["n"] for lev in dev] == [lev["n"] for lev in host]
    for d, h in zip(dev, host):
        assert np.array_equal(d["rowptr"], h["rowptr"]) and np.array_equal(d["col"], h["col"])
        if d["nc"]:
            assert d["nc"] == h["nc"] and np.array_equal(d["Pptr"], h["Pptr"]) and np.array_equal(d["Pcol"], h["Pcol"])
            assert np.array_equal(d["Pval"], h["Pval"])          # same arithmetic, same order, same fp32 rounding


@pytest.mark.parametrize("setup", ["device", "host"])
@pytest.mark.parametrize("which", ["F2", "S"])
def test_device_hierarchy_matches_restatement(D, which, setup, monkeypatch):
    """Hierarchy construction (device-side coarsening of the large levels / host reference path), Galerkin operators
    on every level (k_amg_rap2), their inverse diagonals and one V-cycle (smoother, residual, restriction, coarsest
    inverse, prolongation) against numpy on the same frozen prolongators."""
    if setup == "device":
        monkeypatch.setenv("DDPS_AMG_DEV_MIN", "500")     # small meshes: push them through the device path as well
    else:
        monkeypatch.setenv("DDPS_AMG_HOST_SETUP", "1")
    mesh, pr = MESHES[which](D)
    V, u, v = smooth_fields(mesh, 5)
    with _session(D, mesh, pr) as s:
        s.debug_assemble_base(False)
        sizes = s.amg_levels()
        assert len(sizes) >= 3 and sizes[0] == mesh.nq
        MV = s.get_csr(0)
        levels = s.amg_hierarchy()
        assert [lev["n"] for lev in levels] == sizes
        # level 0 -> 1 sees bitwise identical inputs on the host and on the device: identical aggregates, prolongator and
        # coarse pattern (deeper levels inherit last-bit differences of the coarse base values, so only their sizes and
        # the mathematical properties below are compared)
        host = D.amg.host_hierarchy(MV.indptr, MV.indices, MV.data, _fixed(D, mesh, pr))
        _same_hierarchy(levels[:1], host[:1])
        assert np.array_equal(levels[1]["rowptr"], host[1]["rowptr"]) and np.array_equal(levels[1]["col"], host[1]["col"])
        if setup == "host":
            _same_hierarchy(levels, host)
        s.debug_assemble_uv(1, "shockley", 1.0, 1.0, V, u, v)
        A = s.get_csr(1)
        s.amg_numeric()
        ops, Ps = amg_ref.operator_levels(A, levels)
        for l in range(1, len(levels)):
            # the frozen pattern holds every structural product
            G = ops[l].copy()
            G.data[:] = 1.0
            pat = sp.csr_matrix((np.ones(levels[l]["col"].size), levels[l]["col"], levels[l]["rowptr"]), shape=G.shape)
            assert (G - G.multiply(pat)).nnz == 0 or abs(G - G.multiply(pat)).max() == 0
            val, dinv = s.amg_get_values(l, levels[l]["col"].size)
            dev = sp.csr_matrix((val, levels[l]["col"], levels[l]["rowptr"]), shape=(sizes[l], sizes[l]))
            d = (dev - ops[l]).tocoo()
            err = (np.max(np.abs(d.data)) if d.nnz else 0.0) / np.max(np.abs(ops[l].data))
            assert err <= 1e-12, (l, err)
            assert rel_max(dinv, 1.0 / ops[l].diagonal()) <= 1e-12
        rng = np.random.default_rng(1)
        r = rng.standard_normal(mesh.nq)
        z = s.amg_apply(r)
        zr = amg_ref.vcycle(ops, Ps, r, omega=amg_ref.cycle_omega(MV))
        assert rel_max(z, zr) <= 1e-10
        # run to run bitwise identical
        s.amg_numeric()
        assert np.array_equal(s.amg_apply(r), z)


def test_fp32_cycle_matrices(D, monkeypatch):
    """The large levels keep an fp32 copy of their matrix values for the two SpMVs of the cycle (the cycle is only a
    preconditioner; arithmetic stays FP64).  Forced on a small mesh here and checked against the restatement with the
    same rounding; the solve still reaches the direct solution."""
    monkeypatch.setenv("DDPS_AMG_FP32_MIN", "1000")
    mesh, pr = MESHES["S"](D)
    V, u, v = smooth_fields(mesh, 9)
    with _session(D, mesh, pr) as s:
        s.debug_assemble_base(False)
        om = amg_ref.cycle_omega(s.get_csr(0))
        levels = s.amg_hierarchy()
        assert levels[0]["n"] > 1000 and levels[1]["n"] > 1000
        s.debug_assemble_uv(1, "shockley", 1.0, 1.0, V, u, v)
        A = s.get_csr(1)
        b = s.get_vector(0)
        s.amg_numeric()
        ops, Ps = amg_ref.operator_levels(A, levels)
        r = np.random.default_rng(2).standard_normal(mesh.nq)
        z = s.amg_apply(r)
        assert rel_max(z, amg_ref.vcycle(ops, Ps, r, omega=om, fp32_min=1000)) <= 1e-10
        assert rel_max(z, amg_ref.vcycle(ops, Ps, r, omega=om)) > 1e-12          # it really is the fp32 copy that ran
        x, it, _ = s.debug_pcg(b, rtol=1e-13)
        assert rel_l2(x, spla.spsolve(A.tocsc(), b)) <= 1e-10 and it <= 40


@pytest.mark.parametrize("which", ["F2", "S"])
def test_amg_pcg_against_direct_solve(D, which):
    mesh, pr = MESHES[which](D)
    V, u, v = smooth_fields(mesh, 7)
    with _session(D, mesh, pr) as s:
        s.debug_assemble_base(False)
        s.debug_assemble_uv(1, "shockley", 1.0, 1.0, V, u, v)
        A = s.get_csr(1)
        b = s.get_vector(0)
        x, it, rr = s.debug_pcg(b, rtol=1e-13)
        xd = spla.spsolve(A.tocsc(), b)
        assert rel_l2(x, xd) <= 1e-10
        assert it <= 40, it
        # warm start from the solution: converged at once
        x2, it2, _ = s.debug_pcg(b, x0=x, rtol=1e-12)
        assert it2 <= 1
    with D.Session(0) as sj:   # the Jacobi path on the same system needs an order of magnitude more iterations
        sj.set_mesh(mesh)
        sj.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
        sj.debug_assemble_base(False)
        sj.debug_assemble_uv(1, "shockley", 1.0, 1.0, V, u, v)
        xj, itj, _ = sj.debug_pcg(b, rtol=1e-13)
        assert rel_l2(x, xj) <= 1e-10
        assert itj > 5 * it, (it, itj)


CASES = {
    "ddpe1_F2": ("F2", lambda P: P.ddpe1()),
    "ddpe2_F2": ("F2", lambda P: P.ddpe2()),
    "pn_F1": ("F1", lambda P: P.pn_junction([2, 4], [1, 2, 3, 4], C0=1.0, U=0.3)),
    "pn_F2": ("F2", lambda P: P.pn_junction([3, 7], [1, 2, 3, 4, 5, 6, 7, 8], C0=1.0, U=0.3)),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_solve_ddpe_with_amg_matches_golden(D, case):
    """Same bar as the Jacobi path: fields <= 1e-8 rel-L2, Gummel and Newton iteration counts equal to the oracle's."""
    mname, mk = CASES[case]
    mesh = gpu_mesh(mname)
    z = golden(case)
    pr = mk(D.problems)
    with D.Session(0, precond=1) as s:
        V, u, v, st = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s, return_stats=True, verbose=False)
    assert st["amg_levels"] >= 2
    for a, name in ((V, "V"), (u, "u"), (v, "v")):
        assert rel_l2(a, z[name]) <= FIELD_TOL, (name, rel_l2(a, z[name]))
    # same rule as tests/test_gpu_parity.py: counts equal unless the stop decision sits on the rounding floor
    nref, ngpu = len(z["control"]), st["outer_iters"]
    if ngpu != nref:
        k = min(nref, ngpu) - 1
        assert abs(ngpu - nref) == 1 and pr["tol"] / 4 <= z["control"][k] <= 4 * pr["tol"], (ngpu, nref, z["control"])
    m = min(nref, ngpu)
    assert np.array_equal(st["newton_per_vsolve"].astype(int)[:m], z["newton_counts"][:m])
    assert np.allclose(st["control_history"][:3], z["control"][:3], rtol=1e-4)
    assert max(st["pcg_per_solve"]) <= 40


@pytest.mark.parametrize("tags4", [True, False])
def test_solve_semlin_with_amg_matches_golden(D, tags4):
    mesh = gpu_mesh("F1" if tags4 else "F2")
    z = golden("semlin_F1" if tags4 else "semlin_F2")
    pr = D.problems.semlin_test(tags4)
    with D.Session(0, precond=1) as s:
        V, st = D.solve_semlin_poisson(mesh, pr["A"], pr["bddata"], pr["f"], pr["g"], pr["tol"], session=s, return_stats=True,
                                       verbose=False)
    assert st["amg_levels"] >= 2
    assert rel_l2(V, z["V"]) <= FIELD_TOL
    assert st["outer_iters"] == int(z["newton"])


def test_amg_and_jacobi_agree_on_a_structured_diode_and_renumbering(D, O):
    """Structured pn-diode (the bench problem, 36,864 triangles): multigrid path (with and without the internal Morton
    renumbering) against the CPU oracle at the north-star tolerance, same Gummel / Newton counts as the Jacobi path,
    an order of magnitude fewer CG iterations, bitwise reproducible."""
    mesh = D.structured_mesh(192, 96, 0.0, 2.0, 0.0, 1.0)
    pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
    ref = O.solve_ddpe(oracle_mesh(mesh), *[pr[k] for k in KEYS])
    out = {}
    for name, opts in (("jacobi", {}), ("amg", dict(precond=1)), ("amg_morton", dict(precond=1, renumber=1))):
        with D.Session(0, **opts) as s:
            V, u, v, st = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s, return_stats=True, verbose=False)
        out[name] = (V, u, v, st)
        for a, b in zip((V, u, v), ref):
            assert rel_l2(a, b) <= FIELD_TOL, (name, rel_l2(a, b))
    stj = out["jacobi"][3]
    for name in ("amg", "amg_morton"):
        sta = out[name][3]
        assert sta["outer_iters"] == stj["outer_iters"]
        assert [int(k) for k in sta["newton_per_vsolve"]] == [int(k) for k in stj["newton_per_vsolve"]]
        assert sta["pcg_total"] * 8 < stj["pcg_total"], (sta["pcg_total"], stj["pcg_total"])
    print("rel-L2 vs oracle:", {k: [rel_l2(a, b) for a, b in zip(out[k][:3], ref)] for k in out})
    # bitwise reproducible
    with D.Session(0, precond=1) as s:
        V2, u2, v2 = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s, verbose=False)
    assert np.array_equal(V2, out["amg"][0]) and np.array_equal(u2, out["amg"][1]) and np.array_equal(v2, out["amg"][2])


def test_lagged_coarse_operators(D, O):
    """opts.amg_no_lag = 0 (default): the Galerkin operators of a kind of system (first / later Newton system, u-, v-continuity
    system) are reused while the fine operator stays within 1 % of the one they were computed from.  Only the preconditioner
    changes: same fields (north-star tolerance against the oracle), same Gummel / Newton counts, a few more CG iterations at
    most; the late Gummel iterations do reuse."""
    mesh = D.structured_mesh(160, 80, 0.0, 2.0, 0.0, 1.0)
    pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
    ref = O.solve_ddpe(oracle_mesh(mesh), *[pr[k] for k in KEYS])
    out = {}
    for name, opts in (("lagged", dict(precond=1)), ("exact", dict(precond=1, amg_no_lag=1))):
        with D.Session(0, **opts) as s:
            V, u, v, st = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s, return_stats=True, verbose=False)
        out[name] = (V, u, v, st)
        for a, b in zip((V, u, v), ref):
            assert rel_l2(a, b) <= FIELD_TOL
    sl, se = out["lagged"][3], out["exact"][3]
    assert se["amg_reuses"] == 0 and sl["amg_reuses"] >= 8, (sl["amg_reuses"], se["amg_reuses"])
    assert sl["outer_iters"] == se["outer_iters"]
    assert [int(k) for k in sl["newton_per_vsolve"]] == [int(k) for k in se["newton_per_vsolve"]]
    assert sl["pcg_total"] <= 1.1 * se["pcg_total"] + 4, (sl["pcg_total"], se["pcg_total"])


def test_fused_cycle_tail_equals_per_level_launches(D, monkeypatch):
    """k_amg_tail runs the levels below 32768 rows of the V-cycle in ONE cooperative launch (grid barriers between the phases)
    instead of one launch per level and phase (DDPS_AMG_NO_TAIL=1).  Same arithmetic on every level of <= 4096 rows, a different
    but fixed row-summation order between 4096 and 32768 rows: same Gummel / Newton path, CG counts within one iteration per
    solve, fields equal far below the parity tolerance; and the fused launch is what runs by default (fewer launches)."""
    mesh = D.structured_mesh(320, 160, 0.0, 2.0, 0.0, 1.0)   # levels 51681 / ~8600 / ~960 / ... rows
    pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
    out = {}
    for name in ("fused", "per_level"):
        if name == "per_level":
            monkeypatch.setenv("DDPS_AMG_NO_TAIL", "1")
        with D.Session(0, precond=1) as s:
            V, u, v, st = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s, return_stats=True, verbose=False)
        out[name] = (V, u, v, st)
    monkeypatch.delenv("DDPS_AMG_NO_TAIL")
    sf, sp_ = out["fused"][3], out["per_level"][3]
    assert sf["amg_levels"] >= 4
    assert sf["outer_iters"] == sp_["outer_iters"]
    assert [int(k) for k in sf["newton_per_vsolve"]] == [int(k) for k in sp_["newton_per_vsolve"]]
    assert np.all(np.abs(np.array(sf["pcg_per_solve"]) - np.array(sp_["pcg_per_solve"])) <= 1)
    for a, b in zip(out["fused"][:3], out["per_level"][:3]):
        assert rel_l2(a, b) <= 1e-10, rel_l2(a, b)
    assert sf["kernel_launches"] < 0.8 * sp_["kernel_launches"], (sf["kernel_launches"], sp_["kernel_launches"])


@pytest.mark.parametrize("precond", [0, 1])
def test_warm_started_newton_corrections(D, O, precond):
    """opts.warm_start = 1 (default): the k-th Newton correction of a Vsolve starts its CG from the k-th correction of the previous
    Gummel iteration (the Newton iterates still restart from V = 0, src:311), the continuity solves from the previous iterate.
    Only the CG iteration counts change: same fields (north-star tolerance against the oracle, far closer to each other), same
    Gummel / Newton counts and control history."""
    mesh = D.structured_mesh(160, 80, 0.0, 2.0, 0.0, 1.0)
    pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
    ref = O.solve_ddpe(oracle_mesh(mesh), *[pr[k] for k in KEYS])
    out = {}
    for warm in (1, 0):
        with D.Session(0, precond=precond, warm_start=warm) as s:
            V, u, v, st = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s, return_stats=True, verbose=False)
        out[warm] = (V, u, v, st)
        for a, b in zip((V, u, v), ref):
            assert rel_l2(a, b) <= FIELD_TOL
    sw, sc = out[1][3], out[0][3]
    assert sw["outer_iters"] == sc["outer_iters"]
    assert [int(k) for k in sw["newton_per_vsolve"]] == [int(k) for k in sc["newton_per_vsolve"]]
    assert np.allclose(sw["control_history"], sc["control_history"], rtol=1e-6, atol=1e-10)   # (update norms down to tol = 1e-9; the fields agree to a few 1e-10)
    for a, b in zip(out[1][:3], out[0][:3]):
        assert rel_l2(a, b) <= 5e-9, rel_l2(a, b)
    assert sw["pcg_total"] < 0.85 * sc["pcg_total"], (sw["pcg_total"], sc["pcg_total"])
