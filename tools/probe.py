"""Scale probe (not part of the product): timings at a given structured mesh size."""
import sys, time, json
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import ddps_b200 as D

nx, ny = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
t = time.time(); mesh = D.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0); t_mesh = time.time() - t
pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
s = D.Session(0)
t = time.time(); s.set_mesh(mesh); t_set = time.time() - t
t = time.time(); s.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"]); t_setup = time.time() - t
print(json.dumps(dict(nme=mesh.nme, nq=mesh.nq, t_mesh=t_mesh, t_mesh_set=t_set, t_setup_ddpe=t_setup, lib_setup_ms=s.stats()["t_setup_ms"])), flush=True)
s.ddpe_begin("shockley", 1.0, 1.0, 1.0, 1e-9)
for k in range(steps):
    t = time.time(); c = s.ddpe_step(1); dt = time.time() - t
    st = s.stats()
    print(json.dumps(dict(step=k + 1, control=float(c[0]), wall_s=dt, newton=s.history(1).tolist()[-1:], pcg=s.history(2).tolist()[-8:], t_asm_ms=st["t_asm_ms"], t_pcg_ms=st["t_pcg_ms"], launches=st["kernel_launches"])), flush=True)
for k, name in ((0, "spmv"), (1, "asm_uv"), (2, "pcg_iter"), (3, "asm_newton"), (4, "asm_base")):
    ms, by = s.time_kernel(k, 3, 20)
    print(json.dumps(dict(kernel=name, ms=ms, algo_GB=by / 1e9, GBs=by / ms / 1e6)), flush=True)
