"""Generates the committed fixtures under tests/golden/ (run in the BUILD container, where
/root/reference exists; nothing at test time reads /root/reference).

  python tests/golden/make_golden.py

1. meshes: the reference's own test meshes (test/mesh_p05.msh = F1, test/testmesh.msh = F2) parsed
   with the oracle's MSH reader and stored as compressed .npz (nodes f64, edges/elements i64).
2. golden outputs of the CPU oracle (oracle/ddps_oracle.py) for the reference's test problems
   (test/runtests.jl:14-112) and the builder-defined pn-junction, on both meshes: final fields,
   iteration histories, and the assembled matrices/vectors the parity tests compare against.
   The dense-LU mode (= the reference as written) is run for the semlin problem and stored too.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ddps_oracle as O  # noqa: E402
import ddps_b200.problems as P  # noqa: E402

REF = "/root/reference/test"
KEYS = ("rec_mode", "Vbddata", "ubddata", "vbddata", "lam", "delta", "tau_p", "tau_n", "mu_p", "mu_n", "C", "tol")


def csr_parts(prefix, A):
    A = A.tocsr()
    A.sort_indices()
    return {prefix + "_data": A.data, prefix + "_indices": A.indices.astype(np.int32), prefix + "_indptr": A.indptr.astype(np.int32)}


def four_tag(problem):
    """8-column bddata of the reference's ddpe tests restricted to the 4 tags of F1."""
    out = dict(problem)
    for k in ("Vbddata", "ubddata", "vbddata"):
        tags, kinds, f = problem[k]
        out[k] = [tags[:4], kinds[:4], f[:4]]
    return out


def main():
    meshes = {}
    for name, fn in (("F1", "mesh_p05.msh"), ("F2", "testmesh.msh")):
        m = O.read_mesh(os.path.join(REF, fn))
        meshes[name] = m
        np.savez_compressed(os.path.join(HERE, f"mesh_{name}.npz"), nodes=m.nodes, edges=m.edges, elements=m.elements)
        print(name, m.nodes.shape, m.edges.shape, m.elements.shape)

    # ---- semilinear Poisson (test:93-112) on F1 (as runtests.jl) and F2 (as BASELINE.json words it)
    for name, tags4 in (("F1", True), ("F2", False)):
        pr = P.semlin_test(tags4)
        tr = {}
        V = O.solve_semlin_poisson(meshes[name], pr["A"], pr["bddata"], pr["f"], pr["g"], pr["tol"], trace=tr)
        out = dict(V=V, res=np.array(tr["res"]), newton=tr["newton"], n=tr["n"], gD=tr["gD"], b0=tr["b0"], b_last=tr["b_last"])
        out.update(csr_parts("M", tr["M"]))
        out.update(csr_parts("A_first", tr["A_first"]))
        if name == "F1":
            Vd = O.solve_semlin_poisson(meshes[name], pr["A"], pr["bddata"], pr["f"], pr["g"], pr["tol"], linsolve="dense")
            out["V_dense"] = Vd
            print("  dense vs sparse rel", np.linalg.norm(Vd - V) / np.linalg.norm(V))
        np.savez_compressed(os.path.join(HERE, f"semlin_{name}.npz"), **out)
        print("semlin", name, "newton", tr["newton"], "err", np.max(np.abs(V - pr["exact"](*meshes[name].nodes))))

    # ---- ddpe problems
    cases = {
        "ddpe1_F2": ("F2", P.ddpe1()),
        "ddpe2_F2": ("F2", P.ddpe2()),
        "ddpe1_F1": ("F1", four_tag(P.ddpe1())),
        "ddpe2_F1": ("F1", four_tag(P.ddpe2())),
        "pn_F1": ("F1", P.pn_junction([2, 4], [1, 2, 3, 4], C0=1.0, U=0.3)),
        "pn_F2": ("F2", P.pn_junction([3, 7], [1, 2, 3, 4, 5, 6, 7, 8], C0=1.0, U=0.3)),
    }
    for cname, (mname, pr) in cases.items():
        tr = {}
        V, u, v = O.solve_ddpe(meshes[mname], *[pr[k] for k in KEYS], trace=tr)
        out = dict(V=V, u=u, v=v, control=np.array(tr["control"]), newton_counts=np.array(tr["newton_counts"]),
                   n=tr["n"], gD_V=tr["gD_V"], b_last_u=tr["b_last_u"], b_last_v=tr["b_last_v"])
        out.update(csr_parts("MV", tr["MV"]))
        out.update(csr_parts("A_last_u", tr["A_last_u"]))
        out.update(csr_parts("A_last_v", tr["A_last_v"]))
        np.savez_compressed(os.path.join(HERE, f"{cname}.npz"), **out)
        print(cname, "gummel", len(tr["control"]), "newton", tr["newton_counts"][:6], "control_last", tr["control"][-1])


if __name__ == "__main__":
    main()
