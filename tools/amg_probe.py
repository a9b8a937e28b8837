"""Scale probe of the multigrid-preconditioned path (not part of the product): setup time, Gummel steps, PCG
iterations per solve and kernel timings at a given structured mesh size.
  python tools/amg_probe.py NX NY STEPS [amg|jacobi] [lin_max_it]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddps_b200 as D

nx, ny = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
mode = sys.argv[4] if len(sys.argv) > 4 else "amg"
cap = int(sys.argv[5]) if len(sys.argv) > 5 else 2000
mesh = D.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0)
pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
s = D.Session(0, precond=1 if mode == "amg" else 0, lin_max_it=cap)
t = time.time(); s.set_mesh(mesh); t_set = time.time() - t
s.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
t = time.time(); s.ddpe_begin("shockley", 1.0, 1.0, 1.0, 1e-9); t_begin = time.time() - t
st = s.stats()
print(json.dumps(dict(mode=mode, nme=mesh.nme, nq=mesh.nq, t_mesh_set=t_set, t_ddpe_begin=t_begin, amg_levels=s.amg_levels() if mode == "amg" else [],
                      t_amg_setup_ms=st["t_amg_setup_ms"])), flush=True)
for k in range(steps):
    t = time.time(); c = s.ddpe_step(1); dt = time.time() - t
    st = s.stats()
    print(json.dumps(dict(step=k + 1, control=float(c[0]), wall_s=dt, newton=s.history(1).tolist()[-1:], pcg=s.history(2).tolist()[-9:],
                          t_asm_ms=st["t_asm_ms"], t_pcg_ms=st["t_pcg_ms"], launches=st["kernel_launches"])), flush=True)
kern = ((5, "spmv_dot"), (2, "jacobi_pcg_iter")) + (((6, "amg_pcg_iter"), (8, "vcycle"), (7, "amg_numeric")) if mode == "amg" else ())
for k, name in kern:
    ms, by = s.time_kernel(k, 2, 20)
    print(json.dumps(dict(kernel=name, ms=ms)), flush=True)
