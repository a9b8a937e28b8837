"""A/B timing of the PCG iteration with and without the alternating-traversal trick."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddps_b200 as D
from ddps_b200 import _lib
nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 2048)
mesh = D.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0)
pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
for flag, nosc in ((0, 1), (0, 0)):
    s = D.Session(0, lin_max_it=20, lin_cap_ok=1, max_newton=2)
    o = s._opts; o.no_alt_traversal = flag; o.no_sym_scaling = nosc; s.lib.ddps_set_opts(s.h, __import__("ctypes").byref(o))
    s.set_mesh(mesh)
    s.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
    s.ddpe_begin("shockley", 1.0, 1.0, 1.0, 1e-9)
    s.ddpe_step(1)
    for k, name in ((5, "spmv_dot"), (2, "pcg_iter")):
        ms, by = s.time_kernel(k, 10, 200)
        print(json.dumps(dict(alt_traversal=not flag, sym_scaling=not nosc, kernel=name, ms=ms, GBs=by / ms / 1e6)), flush=True)
    s.close()
