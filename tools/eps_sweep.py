#!/usr/bin/env python
"""How tight must the energy-norm stop rule be?  For DDPS_ETOL_FACTOR in argv: fields of the 512x256 diode against the CPU
oracle (both preconditioners) and, with --s16, the multigrid path at 16.8M triangles against the tightest setting; CG
iterations per Gummel iteration.  Each factor runs in a fresh process (the factor is read once).
  python tools/eps_sweep.py 0.1 0.03 0.01 [--s16]"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
KEYS = ("rec_mode", "Vbddata", "ubddata", "vbddata", "lam", "delta", "tau_p", "tau_n", "mu_p", "mu_n", "C", "tol")

if "--child" in sys.argv:
    import ddps_b200 as D
    pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
    out = {}
    mesh = D.structured_mesh(512, 256, 0.0, 2.0, 0.0, 1.0)
    ref = np.load("/tmp/eps_oracle.npz")
    for name, opts in (("jacobi", {}), ("amg", dict(precond=1))):
        with D.Session(0, **opts) as s:
            V, u, v, st = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s, return_stats=True, verbose=False)
        out[name] = dict(err=[float(np.linalg.norm(a - ref[k]) / np.linalg.norm(ref[k])) for a, k in ((V, "V"), (u, "u"), (v, "v"))],
                         gummel=int(st["outer_iters"]), cg_per_gummel=st["pcg_total"] / st["outer_iters"], newton=[int(x) for x in st["newton_per_vsolve"]])
    if "--s16" in sys.argv:
        mesh = D.structured_mesh(4096, 2048, 0.0, 2.0, 0.0, 1.0)
        with D.Session(0, precond=1) as s:
            V, u, v, st = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s, return_stats=True, verbose=False)
        np.savez(f"/tmp/eps_s16_{os.environ['DDPS_ETOL_FACTOR']}.npz", V=V, u=u, v=v)
        out["s16"] = dict(gummel=int(st["outer_iters"]), cg_per_gummel=st["pcg_total"] / st["outer_iters"], ms=st["t_solve_ms"])
    print("CHILD " + json.dumps(out), flush=True)
    sys.exit(0)

import ddps_b200 as D
import ddps_oracle as O
pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
mesh = D.structured_mesh(512, 256, 0.0, 2.0, 0.0, 1.0)
Vo, uo, vo = O.solve_ddpe(O.Mesh(mesh.nodes, mesh.edges, mesh.elements), *[pr[k] for k in KEYS])
np.savez("/tmp/eps_oracle.npz", V=Vo, u=uo, v=vo)
factors = [a for a in sys.argv[1:] if not a.startswith("--")]
res = {}
for f in factors:
    env = dict(os.environ, DDPS_ETOL_FACTOR=f)
    p = subprocess.run([sys.executable, __file__, "--child"] + [a for a in sys.argv[1:] if a.startswith("--")], env=env, capture_output=True, text=True)
    line = [l for l in p.stdout.splitlines() if l.startswith("CHILD ")]
    res[f] = json.loads(line[0][6:]) if line else dict(error=p.stderr[-400:])
    print(f, json.dumps(res[f]), flush=True)
if "--s16" in sys.argv:
    tight = min(factors, key=float)
    a = np.load(f"/tmp/eps_s16_{tight}.npz")
    for f in factors:
        b = np.load(f"/tmp/eps_s16_{f}.npz")
        print(f"s16 factor {f} vs {tight}:", [float(np.linalg.norm(b[k] - a[k]) / np.linalg.norm(a[k])) for k in "Vuv"])
