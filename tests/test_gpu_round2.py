"""Round-2 GPU parity tests (-m gpu, all through the C ABI): the BIG configurations of BASELINE.json against the CPU
oracle at the same size, both preconditioners of the hand-written CG; device-sampled coefficients (registered spatial
functors) bitwise against host sampling; the opt-in Newton damping against the oracle's restatement of it.

Tolerances (BASELINE.json north_star): final fields <= 1e-8 relative L2, iteration counts equal to the oracle's.
"""
import numpy as np
import pytest

from util import KEYS, gpu_mesh, oracle_mesh, rel_l2

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


# ---------------------------------------------------------------- config 3: solve_semlin_poisson at 1M triangles
def test_config3_semlin_1M_triangles_vs_oracle(D, O):
    """BASELINE.json configs[2] / SURVEY 8d config 3: the test:93-101 problem (SinhRHS functor) on the structured
    724x724-cell mesh (1,048,352 triangles), tol 1e-14, Jacobi- and multigrid-preconditioned CG, against the CPU
    oracle ON THE SAME MESH (about 40 s of sparse LU): fields <= 1e-8, equal Newton counts."""
    mesh = D.structured_mesh(724, 724, -1.0, 1.0, -1.0, 1.0)
    assert mesh.nme == 1048352
    pr = D.problems.semlin_test(True)
    tr = {}
    Vo = O.solve_semlin_poisson(oracle_mesh(mesh), pr["A"], pr["bddata"], pr["f"], pr["g"], pr["tol"], trace=tr)
    for opts in ({}, dict(precond=1)):
        with D.Session(0, **opts) as s:
            V, st = D.solve_semlin_poisson(mesh, pr["A"], pr["bddata"], pr["f"], pr["g"], pr["tol"], session=s,
                                           return_stats=True, verbose=False)
        assert rel_l2(V, Vo) <= FIELD_TOL, (opts, rel_l2(V, Vo))
        assert st["outer_iters"] == tr["newton"], (opts, st["outer_iters"], tr["newton"])
        if opts:
            assert st["amg_levels"] >= 4


# ---------------------------------------------------------------- config 4 scaled to what the oracle can solve
def test_ddpe_diode_262k_triangles_vs_oracle(D, O):
    """The pn-diode of configs[3] on 512x256 cells (262,144 triangles; the largest size the CPU oracle solves in
    under a minute), Gummel to tol 1e-9, BOTH preconditioners: fields <= 1e-8, Gummel and Newton counts equal."""
    mesh = D.structured_mesh(512, 256, 0.0, 2.0, 0.0, 1.0)
    pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
    tr = {}
    Vo, uo, vo = O.solve_ddpe(oracle_mesh(mesh), *[pr[k] for k in KEYS], trace=tr)
    for opts in ({}, dict(precond=1)):
        with D.Session(0, **opts) as s:
            V, u, v, st = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s, return_stats=True, verbose=False)
        for a, b, name in ((V, Vo, "V"), (u, uo, "u"), (v, vo, "v")):
            assert rel_l2(a, b) <= FIELD_TOL, (opts, name, rel_l2(a, b))
        assert st["outer_iters"] == len(tr["control"]), (opts, st["outer_iters"], len(tr["control"]))
        assert [int(k) for k in st["newton_per_vsolve"]] == tr["newton_counts"], opts
        assert np.allclose(st["control_history"], tr["control"], rtol=1e-2, atol=1e-11), opts


def test_config4_16M_triangles_amg_vs_jacobi_to_tol(D):
    """configs[3] at FULL size (4096x2048 cells, 16,777,216 triangles), Gummel to tol 1e-9 with each preconditioner:
    two independent solvers (Jacobi-PCG needs ~23,000 CG iterations per Gummel iteration, the multigrid path ~70) must
    agree to 1e-8 and take the same Gummel/Newton path -- the heuristics of the multigrid path (lagged coarse operators,
    energy-norm stop rule) are exercised at 8.4M rows."""
    mesh = D.structured_mesh(4096, 2048, 0.0, 2.0, 0.0, 1.0)
    pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
    res = {}
    for name, opts in (("amg", dict(precond=1)), ("jacobi", {})):
        with D.Session(0, **opts) as s:
            V, u, v, st = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s, return_stats=True, verbose=False)
        res[name] = (V, u, v, st)
    for k, fname in enumerate("Vuv"):
        assert rel_l2(res["amg"][k], res["jacobi"][k]) <= FIELD_TOL, (fname, rel_l2(res["amg"][k], res["jacobi"][k]))
    sa, sj = res["amg"][3], res["jacobi"][3]
    assert sa["outer_iters"] == sj["outer_iters"]
    assert list(sa["newton_per_vsolve"]) == list(sj["newton_per_vsolve"])
    assert sa["pcg_total"] * 50 < sj["pcg_total"]
    # contact values are reproduced (Dirichlet rows are identity rows, src:279-284)
    n0 = D.dirichlet_nodes(mesh, pr["Vbddata"])[0] - 1
    gV = pr["Vbddata"][2][1](mesh.nodes[0, n0], mesh.nodes[1, n0])
    assert np.allclose(res["amg"][0][n0], gV, rtol=0, atol=1e-9)


# ---------------------------------------------------------------- device-side coefficient sampling
@pytest.mark.parametrize("which", ["F2", "S"])
def test_device_sampled_coefficients_equal_host_sampling_bitwise(D, which):
    """Registered spatial functors are evaluated by the library on the device at the reference's sampling points
    (centroids src:250-254,463-467; edge midpoints src:303-308): the arrays equal host sampling of the same closures
    BITWISE, so switching between the two cannot change any result."""
    from ddps_b200 import api
    mesh = gpu_mesh("F2") if which == "F2" else D.structured_mesh(96, 40, 0.0, 2.0, 0.0, 1.0)
    fns = [D.Const(1.37), D.Affine(0.4, 0.25, -0.15), D.StepX(0.13, -2.0, 3.0), D.StepY(0.31, 0.5, 1.5)]
    cx, cy = api._centroids(mesh)
    mx, my = api._midpoints(mesh)
    with D.Session(0) as s:
        s.set_mesh(mesh)
        for f in fns:
            s.set_coeff_functor(api.COEFF_LAMBDA2, f)
            assert np.array_equal(s.get_coeff(api.COEFF_LAMBDA2), api._sample2(f, cx, cy) ** 2)
            s.set_coeff_functor(api.COEFF_MU_P, f)
            assert np.array_equal(s.get_coeff(api.COEFF_MU_P), api._sample2(f, cx, cy))
            s.set_coeff_functor(api.COEFF_C_MID, f)
            assert np.array_equal(s.get_coeff(api.COEFF_C_MID), np.ascontiguousarray(api._sample2(f, mx, my)).reshape(-1))
        with pytest.raises(D.DDPSError):
            p = np.array([1.0, 2.0])
            s._check(s.lib.ddps_coeff_functor(s.h, api.COEFF_MU_P, 3, p.ctypes.data_as(D._lib.c_double_p), 2))


def test_solve_ddpe_device_functors_equal_host_sampled_closures(D):
    """solve_ddpe with registered spatial functors (sampled on the device) == the same solve with plain lambdas (sampled
    on the host, in threads): identical fields, bit for bit."""
    mesh = D.structured_mesh(96, 48, 0.0, 2.0, 0.0, 1.0)
    pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
    plain = dict(pr)
    for k in ("lam", "mu_p", "mu_n", "C"):
        f = pr[k]
        plain[k] = (lambda g: (lambda x, y: g(x, y)))(f)      # hide the functor type: forces host sampling
    a = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], verbose=False)
    b = D.solve_ddpe(mesh, *[plain[k] for k in KEYS], verbose=False)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


# ---------------------------------------------------------------- opt-in Newton damping
def test_newton_damping_opt_in(D, O):
    """opts.newton_damping (default 0 = the reference's undamped update, src:1063-1064): on a semilinear problem whose
    full Newton steps overshoot (strong source, exp-type nonlinearity) the undamped iteration matches the oracle's count
    (14), the damped one matches the oracle's restatement of the same backtracking rule (8) -- same solution."""
    from ddps_b200.functors import SinhRHS
    mesh = gpu_mesh("F1")
    om = oracle_mesh("F1")
    pr = D.problems.semlin_test(True)
    f = SinhRHS(lambda x, y: 50.0 * np.exp(-x ** 2 - y ** 2), alpha=-1.0, beta=1.0)
    out = {}
    for damp in (0, 8):
        tr = {}
        Vo = O.solve_semlin_poisson(om, pr["A"], pr["bddata"], f, f.derivative(), 1e-12, trace=tr, damping=damp)
        with D.Session(0, newton_damping=damp) as s:
            V, st = D.solve_semlin_poisson(mesh, pr["A"], pr["bddata"], f, f.derivative(), 1e-12, session=s,
                                           return_stats=True, verbose=False)
        assert rel_l2(V, Vo) <= FIELD_TOL
        assert st["outer_iters"] == tr["newton"], (damp, st["outer_iters"], tr["newton"])
        assert (st["newton_halvings"] > 0) == (damp > 0)
        out[damp] = (V, st["outer_iters"])
    assert out[8][1] < out[0][1]
    assert rel_l2(out[8][0], out[0][0]) <= FIELD_TOL
