"""Node-partitioned solves with the ranks running as threads on ONE GPU (in-process communicator): N-rank results against the
single-rank result and the golden fixtures, both preconditioners, both transports of the hot exchanges (staged peer-memory
kernels with plain pointers / host-ordered backend exchange).  CUDA IPC and NCCL run with one process per GPU in
tests/test_multigpu.py, which repeats these checks."""
import numpy as np
import pytest

from localranks import run_ranks
from util import KEYS, golden, gpu_mesh, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    import ddps_b200
    return ddps_b200


def _solve_ddpe(D, mesh, pr, comm=None, **opts):
    with D.Session(0, comm=comm, **opts) as s:
        V, u, v, st = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s, return_stats=True, verbose=False)
    return V, u, v, st


@pytest.mark.parametrize("p2p", ["1", "0"])
@pytest.mark.parametrize("nranks", [2, 3])
def test_jacobi_path_local_ranks(D, nranks, p2p, monkeypatch):
    monkeypatch.setenv("DDPS_P2P", p2p)
    mesh = D.structured_mesh(96, 48, 0.0, 2.0, 0.0, 1.0)
    pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0)
    ref = _solve_ddpe(D, mesh, pr)
    res = run_ranks(nranks, lambda r, comm: _solve_ddpe(D, mesh, pr, comm=comm))
    for V, u, v, st in res:
        for a, b in ((V, ref[0]), (u, ref[1]), (v, ref[2])):
            assert rel_l2(a, b) <= 1e-10, rel_l2(a, b)
        assert st["outer_iters"] == ref[3]["outer_iters"]
        assert list(st["newton_per_vsolve"]) == list(ref[3]["newton_per_vsolve"])
    # every rank returns the same full fields
    assert np.array_equal(res[0][0], res[-1][0])


def test_semlin_and_golden_local_ranks(D):
    mesh = gpu_mesh("F1")
    pr = D.problems.semlin_test(True)

    def semlin(r, comm):
        with D.Session(0, comm=comm) as s:
            return D.solve_semlin_poisson(mesh, pr["A"], pr["bddata"], pr["f"], pr["g"], pr["tol"], session=s, return_stats=True, verbose=False)

    z = golden("semlin_F1")
    for V, st in run_ranks(3, semlin):
        assert rel_l2(V, z["V"]) <= 1e-8
        assert st["outer_iters"] == 5
    mesh2 = gpu_mesh("F2")
    pr2 = D.problems.ddpe2()
    z = golden("ddpe2_F2")
    for V, u, v, st in run_ranks(2, lambda r, comm: _solve_ddpe(D, mesh2, pr2, comm=comm)):
        for a, k in ((V, "V"), (u, "u"), (v, "v")):
            assert rel_l2(a, z[k]) <= 1e-8, (k, rel_l2(a, z[k]))
        assert st["outer_iters"] == len(z["control"])


@pytest.mark.parametrize("p2p", ["1", "0"])
@pytest.mark.parametrize("nranks,tail,cells", [(2, 1100, (96, 48)), (3, 1100, (128, 64)), (2, 200000, (96, 48)), (4, 1100, (192, 96))])
def test_distributed_multigrid_local_ranks(D, nranks, tail, cells, p2p, monkeypatch):
    """The distributed multigrid hierarchy (rank-local aggregates, ghost prolongator / A*P rows, halo exchanges of the cycle
    vectors, replicated tail below DDPS_AMG_TAIL rows) against the single-GPU hierarchy on the same problem: fields <= 1e-10,
    the same Gummel and Newton counts, and at most 3 more CG iterations in any linear solve."""
    monkeypatch.setenv("DDPS_P2P", p2p)
    monkeypatch.setenv("DDPS_AMG_TAIL", str(tail))
    mesh = D.structured_mesh(cells[0], cells[1], 0.0, 2.0, 0.0, 1.0)
    pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0)
    ref = _solve_ddpe(D, mesh, pr, precond=1)
    assert ref[3]["amg_levels"] >= 2
    res = run_ranks(nranks, lambda r, comm: _solve_ddpe(D, mesh, pr, comm=comm, precond=1))
    for V, u, v, st in res:
        assert st["amg_levels"] >= 2, "distributed hierarchy was not built (fell back to Jacobi-PCG)"
        for a, b in ((V, ref[0]), (u, ref[1]), (v, ref[2])):
            assert rel_l2(a, b) <= 1e-10, rel_l2(a, b)
        assert st["outer_iters"] == ref[3]["outer_iters"]
        assert list(st["newton_per_vsolve"]) == list(ref[3]["newton_per_vsolve"])
        its, its_ref = np.array(st["pcg_per_solve"]), np.array(ref[3]["pcg_per_solve"])
        assert its.shape == its_ref.shape
        assert np.all(its <= its_ref + 3), (its.tolist(), its_ref.tolist())
    assert np.array_equal(res[0][0], res[-1][0])


def test_distributed_multigrid_semlin_local_ranks(D, monkeypatch):
    monkeypatch.setenv("DDPS_AMG_TAIL", "400")
    mesh = gpu_mesh("F1")
    pr = D.problems.semlin_test(True)

    def semlin(r, comm):
        with D.Session(0, comm=comm, precond=1) as s:
            return D.solve_semlin_poisson(mesh, pr["A"], pr["bddata"], pr["f"], pr["g"], pr["tol"], session=s, return_stats=True, verbose=False)

    z = golden("semlin_F1")
    for V, st in run_ranks(3, semlin):
        assert st["amg_levels"] >= 2
        assert rel_l2(V, z["V"]) <= 1e-8
        assert st["outer_iters"] == 5


@pytest.mark.parametrize("precond", [0, 1])
def test_per_rank_mesh_ingest_local_ranks(D, precond, monkeypatch):
    """ddps_mesh_set_local: every rank passes only its strip of the structured mesh (nobody holds the global mesh) and
    fetches only the nodes it owns; the assembled fields equal the single-GPU solve of the global mesh."""
    monkeypatch.setenv("DDPS_AMG_TAIL", "500")
    nx, ny, nr = 96, 48, 3
    pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0)
    mesh = D.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0)
    ref = _solve_ddpe(D, mesh, pr, precond=precond)

    def work(r, comm):
        part = D.structured_mesh_part(nx, ny, 0.0, 2.0, 0.0, 1.0, r, nr)
        with D.Session(0, comm=comm, precond=precond) as s:
            s.set_mesh_local(part, with_elem_glob=(r % 2 == 0))
            s.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
            s.ddpe_begin(pr["rec_mode"], pr["delta"], pr["tau_p"], pr["tau_n"], pr["tol"])
            ctrl = []
            while not ctrl or ctrl[-1] > pr["tol"]:
                ctrl += list(s.ddpe_step(1, stop_at_tol=True))
            V, u, v = s.ddpe_fetch_owned()
            return part["own_glob"], V, u, v, len(ctrl), s.stats()

    out = run_ranks(nr, work)
    V, u, v = np.zeros(mesh.nq), np.zeros(mesh.nq), np.zeros(mesh.nq)
    for g, Vr, ur, vr, nit, st in out:
        V[g], u[g], v[g] = Vr, ur, vr
        assert nit == ref[3]["outer_iters"]
        assert (st["amg_levels"] >= 2) == bool(precond)
    for a, b in ((V, ref[0]), (u, ref[1]), (v, ref[2])):
        assert rel_l2(a, b) <= 1e-10, rel_l2(a, b)
