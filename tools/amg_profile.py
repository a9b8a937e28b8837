"""Short multigrid profiling driver for ncu (a few hundred launches): S16 pn-diode, one Gummel step with every CG solve
capped at 3 iterations (so the fields are non-trivial), then the per-operator numeric phase and two full
multigrid-preconditioned CG iterations through ddps_time_kernel."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddps_b200 as D

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 2048)
mesh = D.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0)
pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
s = D.Session(0, precond=1, lin_max_it=3, lin_cap_ok=1, max_newton=1)
s.set_mesh(mesh)
s.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
s.ddpe_begin("shockley", 1.0, 1.0, 1.0, 1e-9)
s.ddpe_step(1)
for k, name in ((7, "amg_numeric"), (6, "amg_pcg_iter")):
    ms, by = s.time_kernel(k, 1, 2)
    print(json.dumps(dict(kernel=name, ms=ms)), flush=True)
