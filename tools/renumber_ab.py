"""SpMV / PCG-iteration timing on a randomly numbered mesh with and without the internal Morton renumbering."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddps_b200 as D
nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 1024)
base = D.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0)
rng = np.random.default_rng(0)
p = rng.permutation(base.nq)
nodes = np.empty_like(base.nodes); nodes[:, p] = base.nodes
el = base.elements.copy(); el[0:3] = p[base.elements[0:3] - 1] + 1
ed = base.edges.copy(); ed[0:2] = p[base.edges[0:2] - 1] + 1
shuffled = D.Mesh(nodes, ed, el)
pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
for name, mesh, ren in (("lexicographic", base, 0), ("random numbering", shuffled, 0), ("random + morton renumber", shuffled, 1),
                        ("lexicographic + morton renumber", base, 1)):
    s = D.Session(0, lin_max_it=20, lin_cap_ok=1, max_newton=2, renumber=ren)
    s.set_mesh(mesh)
    s.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
    s.ddpe_begin("shockley", 1.0, 1.0, 1.0, 1e-9)
    s.ddpe_step(1)
    out = dict(mesh=name, triangles=mesh.nme, setup_ms=s.stats()["t_setup_ms"])
    for k, kn in ((5, "spmv_dot"), (2, "pcg_iter"), (1, "asm_uv")):
        ms, by = s.time_kernel(k, 5, 50)
        out[kn + "_ms"] = round(ms, 4); out[kn + "_GBs"] = round(by / ms / 1e6, 1)
    print(json.dumps(out), flush=True)
    s.close()
