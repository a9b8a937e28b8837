"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel of the path on tiny meshes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ddps_b200 as D
from util import KEYS, fan_mesh

mesh = fan_mesh(14, 3)
f = D.SinhRHS(lambda x, y: 1.0 + x * y, alpha=-1.0, beta=1.0)
A = [lambda t: 1.0, lambda t: 0.0, lambda t: 0.0, lambda t: 2.0]
gd = lambda x, y: 0.2 * x - 0.1 * y
bd = [[1, 2], ["D", "N"], [gd, lambda x, y: 0.3 + 0 * x]]
V = D.solve_semlin_poisson(mesh, A, bd, f, f.derivative(), 1e-12, verbose=False)
print("semlin ok", float(np.linalg.norm(V)))
mesh = D.structured_mesh(37, 19, 0.0, 2.0, 0.0, 1.0)   # sizes that are not multiples of 32/128
pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0)
for mode in ("shockley", "zero"):
    p2 = dict(pr, rec_mode=mode)
    Vg, ug, vg = D.solve_ddpe(mesh, *[p2[k] for k in KEYS], verbose=False)
    print("ddpe", mode, "ok", float(np.linalg.norm(ug)))
with D.Session(0) as s:
    s.set_mesh(mesh)
    s.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
    s.ddpe_begin("shockley", 1.0, 1.0, 1.0, 1e-9)
    s.ddpe_step(1)
    for k in (0, 1, 2, 3, 4, 5):
        s.time_kernel(k, 1, 2)
    A_ = s.get_csr(1)
print("done", A_.nnz)
