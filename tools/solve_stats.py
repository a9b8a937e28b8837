"""Where a whole S16 solve spends its time: library statistics of one resident solve_ddpe to tol (scratch helper)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddps_b200 as D

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 2048)
mesh = D.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0)
pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
s = D.Session(0, precond=1)
s.set_mesh(mesh)
s.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
for rep in range(2):
    t0 = time.perf_counter()
    s.ddpe_begin("shockley", 1.0, 1.0, 1.0, 1e-9)
    t1 = time.perf_counter()
    c = s.ddpe_step(1000, stop_at_tol=True)
    t2 = time.perf_counter()
    st = s.stats()
    print(json.dumps(dict(begin_s=round(t1 - t0, 4), steps_s=round(t2 - t1, 4), iters=len(c),
                          **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items() if not hasattr(v, "__len__")})), flush=True)
    print("pcg per solve", s.history(2).astype(int).tolist(), flush=True)
