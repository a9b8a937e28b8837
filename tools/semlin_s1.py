"""solve_semlin_poisson (test:93-101 problem, SinhRHS functor) on a structured 724x724-cell
triangulation of [-1,1]^2 (1,048,352 triangles), one B200: Newton iterations/s with the Jacobi- and the
multigrid-preconditioned CG, agreement of the two, error against the analytic solution and (--oracle) against the CPU oracle.
  python tools/semlin_s1.py [N] [tol] [--oracle]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ddps_b200 as D

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if args else 724
tol = float(args[1]) if len(args) > 1 else 1e-14
mesh = D.structured_mesh(n, n, -1.0, 1.0, -1.0, 1.0)
pr = D.problems.semlin_test(True)
out = {}
for name, opts in (("jacobi", {}), ("amg", dict(precond=1))):
    with D.Session(0, **opts) as s:
        s.set_mesh(mesh)
        s.setup_semlin(pr["A"], pr["bddata"], pr["f"], pr["g"])
        t0 = time.perf_counter()
        V, st = s.solve_semlin(tol)
        wall = time.perf_counter() - t0
        hist = s.history(0)
        pcg = s.history(2)
    out[name] = V
    print(json.dumps(dict(solver=name, triangles=mesh.nme, nodes=mesh.nq, tol=tol, newton_iters=st["outer_iters"], wall_s=wall,
                          newton_iters_per_s=st["outer_iters"] / wall, pcg_per_solve=[int(k) for k in pcg], t_pcg_ms=st["t_pcg_ms"],
                          t_asm_ms=st["t_asm_ms"], t_amg_setup_ms=st["t_amg_setup_ms"], residuals=[float(r) for r in hist],
                          max_err_vs_exact=float(np.max(np.abs(V - pr["exact"](*mesh.nodes)))))), flush=True)
d = np.linalg.norm(out["amg"] - out["jacobi"]) / np.linalg.norm(out["jacobi"])
print(json.dumps(dict(rel_l2_amg_vs_jacobi=float(d))), flush=True)
if "--oracle" in sys.argv:
    import ddps_oracle as O
    t0 = time.perf_counter()
    Vo = O.solve_semlin_poisson(O.Mesh(mesh.nodes, mesh.edges, mesh.elements), pr["A"], pr["bddata"], pr["f"], pr["g"], tol)
    print(json.dumps(dict(oracle_wall_s=time.perf_counter() - t0,
                          rel_l2_jacobi_vs_oracle=float(np.linalg.norm(out["jacobi"] - Vo) / np.linalg.norm(Vo)),
                          rel_l2_amg_vs_oracle=float(np.linalg.norm(out["amg"] - Vo) / np.linalg.norm(Vo)))), flush=True)
