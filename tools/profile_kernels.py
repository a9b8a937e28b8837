"""Short kernel-profiling driver (few hundred launches) for ncu: sets up the S16 pn-diode, runs one
Gummel iteration with every PCG solve capped at 20 iterations (so the fields are non-trivial), then
launches each hot kernel a few times through ddps_time_kernel."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddps_b200 as D

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 2048)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
which = sys.argv[4].split(",") if len(sys.argv) > 4 else None
mesh = D.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0)
pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
s = D.Session(0, lin_max_it=20, lin_cap_ok=1, max_newton=2)
s.set_mesh(mesh)
s.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
s.ddpe_begin("shockley", 1.0, 1.0, 1.0, 1e-9)
s.ddpe_step(1)
for k, name in ((5, "spmv_dot"), (2, "pcg_iter"), (1, "asm_uv"), (3, "asm_newton"), (4, "asm_base"), (14, "prep_poisson"),
                (15, "prep_uv")):
    if which and name not in which:
        continue
    ms, by = s.time_kernel(k, 1, reps)
    print(json.dumps(dict(kernel=name, ms=ms, algo_GB=by / 1e9, GBs=by / ms / 1e6)), flush=True)
