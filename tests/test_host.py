"""CPU tests of the host-side mirror (no GPU): boundary preprocessing against the oracle, mesh
reader/writers, functors, and that the C-ABI library loads and exports every declared symbol."""
import ctypes
import os
import re

import numpy as np
import pytest

from util import four_tag, gpu_mesh, oracle_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def D():
    import ddps_b200
    return ddps_b200


@pytest.fixture(scope="module")
def O():
    import ddps_oracle
    return ddps_oracle


def test_library_exports_every_declared_symbol(D):
    from ddps_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ddps.h")).read()
    declared = set(re.findall(r"\b(ddps_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"ddps_ctx"}
    lib = _lib.load()
    assert lib.ddps_version() == 100
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in ddps.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_no_cpu_fallback(D):
    """Without a CUDA device the product path must fail loudly (there is no CPU fallback)."""
    from ddps_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.ddps_ctx_create(ctypes.byref(h), 0)
    if rc == 0:
        lib.ddps_ctx_destroy(h)
        pytest.skip("a GPU is present")
    assert rc < 0 and b"no CPU fallback" in lib.ddps_last_error(None)
    with pytest.raises(D.DDPSError):
        D.Session(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "driftdiffusionpoissonsystems.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "ddps_oracle" not in src and "import oracle" not in src, f


@pytest.mark.parametrize("mname", ["F1", "F2"])
def test_boundary_preprocessing_matches_oracle(D, O, mname):
    mesh, om = gpu_mesh(mname), oracle_mesh(mname)
    for pr in (D.problems.ddpe1(), D.problems.ddpe2()):
        if mname == "F1":
            pr = four_tag(pr)
        for key in ("Vbddata", "ubddata", "vbddata"):
            n1, n2, gD, n = D.aquire_boundary(mesh, pr[key])
            o1, o2, ogD, on = O.aquire_boundary(om, pr[key])
            assert np.array_equal(n, on) and np.array_equal(gD, ogD)
            assert np.array_equal(n1, o1) and np.array_equal(n2, o2)
    sp = D.problems.semlin_test(mname == "F1")
    a, b, load = D.neumann_edges(mesh, sp["bddata"])
    oa, ob, oload = O.neumann_data(om, sp["bddata"])
    assert np.array_equal(a, oa) and np.array_equal(b, ob) and np.allclose(load, oload, rtol=0, atol=0)


def test_bddata_errors(D):
    mesh = gpu_mesh("F1")
    f = lambda x, y: 0.0 * x
    with pytest.raises(ValueError):      # src:169-171
        D.aquire_boundary(mesh, [[1, 2, 3, 4], ["D", "?", "D", "N"], [f] * 4])
    with pytest.raises(ValueError):
        D.aquire_boundary(mesh, [[1, 2, 3], ["D", "N"], [f] * 3])


def test_read_mesh_roundtrip(D, tmp_path):
    mesh = gpu_mesh("F1")
    path = tmp_path / "m.msh"
    with open(path, "w") as fh:
        fh.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % mesh.nq)
        for i in range(mesh.nq):
            fh.write("%d %.17g %.17g 0\n" % (i + 1, mesh.nodes[0, i], mesh.nodes[1, i]))
        fh.write("$EndNodes\n$Elements\n%d\n" % (mesh.edges.shape[1] + mesh.nme))
        k = 1
        for e in mesh.edges.T:
            fh.write("%d 1 2 %d %d %d %d\n" % (k, e[2], e[3], e[0], e[1]))
            k += 1
        for t in mesh.elements.T:
            fh.write("%d 2 2 %d %d %d %d %d\n" % (k, t[3], t[4], t[0], t[1], t[2]))
            k += 1
        fh.write("$EndElements\n")
    m2 = D.read_mesh(str(path))
    assert np.array_equal(m2.nodes, mesh.nodes) and np.array_equal(m2.edges, mesh.edges)
    assert np.array_equal(m2.elements, mesh.elements)
    # trailing sections warn (src:149-151); other element types error (src:140-142)
    with open(path, "a") as fh:
        fh.write("$Periodic\n0\n$EndPeriodic\n")
    with pytest.warns(UserWarning):
        D.read_mesh(str(path))
    bad = tmp_path / "bad.msh"
    bad.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n1 0 0 0\n$EndNodes\n$Elements\n1\n1 15 2 0 1 1\n$EndElements\n")
    with pytest.raises(ValueError):
        D.read_mesh(str(bad))


def test_structured_mesh_matches_oracle_generator(D, O):
    a, b = D.structured_mesh(7, 5, 0, 2, 0, 1), O.structured_mesh(7, 5, 0, 2, 0, 1)
    assert np.allclose(a.nodes, b.nodes, atol=1e-15) and np.array_equal(a.edges, b.edges)
    assert np.array_equal(a.elements, b.elements)
    assert np.all(O._geometry(b)[6] > 0)


def test_construct_mesh_writers(D, tmp_path):
    E = np.array([[-1.0, -1.0], [1.0, -1.0], [1.0, 1.0], [-1.0, 1.0]])
    D.construct_mesh8(E, np.array([[-0.1, 0.1], [-0.1, 0.1]]), 0.05, str(tmp_path / "t8"))
    txt = (tmp_path / "t8.geo").read_text()
    assert "Point(3)={1.0, -0.1, 0, h};" in txt and "Point(7)={-1.0, 0.1, 0, h};" in txt     # test/testmesh.geo
    assert "Line(8) = {8, 1};" in txt and "Line Loop(1) = {1,2,3,4,5,6,7,8};" in txt
    D.construct_mesh4(E, 0.05, str(tmp_path / "t4"))
    assert "Line(4) = {4, 1};" in (tmp_path / "t4.geo").read_text()
    assert D.construct_mesh is D.construct_mesh8


def test_functors(D):
    f = D.SinhRHS(lambda x, y: x + y, alpha=-1.0, beta=2.0)
    x, y, u = np.array([0.1, 0.2]), np.array([0.3, 0.4]), np.array([0.5, -0.5])
    assert np.allclose(f(x, y, u), x + y - np.sinh(2 * u))
    h = 1e-6
    assert np.allclose(f.derivative()(x, y, u), (f(x, y, u + h) - f(x, y, u - h)) / (2 * h), atol=1e-8)
    p = D.PolyRHS([lambda x, y: 1 + x, lambda x, y: y, lambda x, y: 4 * x * y])
    assert np.allclose(p(x, y, u), 1 + x + y * u + 4 * x * y * u ** 2)
    assert np.allclose(p.derivative()(x, y, u), y + 8 * x * y * u)
    with pytest.raises(ValueError):
        D.PolyRHS([lambda x, y: 0] * 7)


def test_partition_rcb(D):
    """ddps_partition_nodes (host-only entry point): balanced, deterministic, spatially compact."""
    from ddps_b200 import _lib
    lib = _lib.load()
    mesh = D.structured_mesh(40, 20, 0, 2, 0, 1)
    nodes = np.ascontiguousarray(mesh.nodes.T)
    for parts in (1, 2, 4, 8, 3):
        owner = np.full(mesh.nq, -1, dtype=np.int32)
        rc = lib.ddps_partition_nodes(_lib.dptr(nodes), mesh.nq, parts, owner.ctypes.data_as(_lib.c_int32_p))
        assert rc == 0 and owner.min() == 0 and owner.max() == parts - 1
        counts = np.bincount(owner, minlength=parts)
        assert counts.max() - counts.min() <= parts
        o2 = np.full(mesh.nq, -1, dtype=np.int32)
        lib.ddps_partition_nodes(_lib.dptr(nodes), mesh.nq, parts, o2.ctypes.data_as(_lib.c_int32_p))
        assert np.array_equal(owner, o2)
    # 2 parts of a 2:1 rectangle split along x
    owner = np.zeros(mesh.nq, dtype=np.int32)
    lib.ddps_partition_nodes(_lib.dptr(nodes), mesh.nq, 2, owner.ctypes.data_as(_lib.c_int32_p))
    assert mesh.nodes[0, owner == 0].max() <= mesh.nodes[0, owner == 1].min() + 1e-12


def test_terminal_current_postprocessing_matches_oracle(D, O):
    """aquire_boundary_separate_edges / get_gradients / the flux integrals of calculate_current
    (src:662-887) on the host, against the literal oracle restatement (same fields in, same current out)."""
    mesh, om = gpu_mesh("F2"), oracle_mesh("F2")
    E = np.array([[-1.0, -1.0], [1.0, -1.0], [1.0, 1.0], [-1.0, 1.0]])
    pr = D.problems.pn_junction([3, 7], list(range(1, 9)), C0=1.0, U=0.3)
    z = np.load(os.path.join(ROOT, "tests", "golden", "pn_F2.npz"))
    got = D.aquire_boundary_separate_edges(mesh, E, pr["Vbddata"])
    ref = O.aquire_boundary_separate_edges(om, E, pr["Vbddata"])
    for a, b in zip(got, ref):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    assert got[4].shape[1] == 4 and got[6].shape[1] == 4 and got[5].shape[1] == 0      # contacts: 4 edges left, 4 right
    I = D.current_from_fields(mesh, E, pr["Vbddata"], pr["mu_p"], pr["mu_n"], z["u"], z["v"])
    Iref = O.calculate_current(om, pr["rec_mode"], E, pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["delta"],
                               pr["tau_p"], pr["tau_n"], pr["mu_p"], pr["mu_n"], pr["C"], pr["tol"],
                               fields=(z["V"], z["u"], z["v"]))[0]
    assert abs(I - Iref) <= 1e-12 * max(1.0, abs(Iref)) and I != 0.0
    # all-Dirichlet problem: every side contributes
    pr1 = D.problems.ddpe1()
    z1 = np.load(os.path.join(ROOT, "tests", "golden", "ddpe1_F2.npz"))
    I1 = D.current_from_fields(mesh, E, pr1["Vbddata"], pr1["mu_p"], pr1["mu_n"], z1["u"], z1["v"])
    I1ref = O.calculate_current(om, "zero", E, pr1["Vbddata"], pr1["ubddata"], pr1["vbddata"], pr1["lam"], 1.0, 0.0, 0.0,
                                pr1["mu_p"], pr1["mu_n"], pr1["C"], 1e-12, fields=(z1["V"], z1["u"], z1["v"]))[0]
    assert abs(I1 - I1ref) <= 1e-11 * max(1.0, abs(I1ref))


def test_structured_mesh_parts_tile_the_global_mesh():
    """Per-rank mesh parts (ddps_mesh_set_local input): owned nodes partition the nodes, every triangle touching an owned
    node is present in increasing global id with consistent local ids, ghosts are exactly the other corners of those
    triangles, coordinates equal the global mesh's bit for bit, and the boundary-only mesh gives the same Dirichlet lists."""
    import ddps_b200 as D
    nx, ny = 12, 5
    m = D.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0)
    for nr in (1, 2, 3, 4):
        seen = np.zeros(m.nq, int)
        for r in range(nr):
            p = D.structured_mesh_part(nx, ny, 0.0, 2.0, 0.0, 1.0, r, nr)
            g = p["node_glob"] - 1
            assert np.array_equal(p["nodes"][:, 0], m.nodes[0, g]) and np.array_equal(p["nodes"][:, 1], m.nodes[1, g])
            assert np.all(np.diff(g[:p["n_own"]]) > 0)
            seen[g[:p["n_own"]]] += 1
            eg = p["elem_glob"] - 1
            assert np.array_equal(g[p["elements"][:, :3] - 1] + 1, m.elements[:3, eg].T)
            owned = np.zeros(m.nq, bool)
            owned[g[:p["n_own"]]] = True
            assert np.array_equal(np.nonzero(owned[m.elements[:3] - 1].any(axis=0))[0], eg)
            assert set(np.unique(m.elements[:3, eg] - 1).tolist()) == set(g.tolist())
            if nr > 1:
                go = p["ghost_owner"]
                key = go.astype(np.int64) * m.nq + g[p["n_own"]:]
                assert np.all(np.diff(key) > 0)
        assert np.all(seen == 1)
        bd = [[2, 4], ["D", "D"], [lambda x, y: x + 2 * y] * 2]
        n1, v1 = D.dirichlet_nodes(p["boundary"], bd)
        n2, v2 = D.dirichlet_nodes(m, bd)
        assert np.array_equal(n1, n2) and np.array_equal(v1, v2)
