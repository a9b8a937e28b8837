"""Where the end-to-end (cold start) time of bench.py goes: per-phase host wall clock of the same call sequence."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddps_b200 as D

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 2048)
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)
mesh = D.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0)
smp = D.api.sample_ddpe(mesh, pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
nodes = torch.from_numpy(np.ascontiguousarray(mesh.nodes.T)).pin_memory().numpy()
elems = torch.from_numpy(np.ascontiguousarray(mesh.elements.T)).pin_memory().numpy()
mesh = D.Mesh(nodes.T, mesh.edges, elems.T)   # zero-copy view of pinned buffers in the ABI layout
pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in smp.items()}
torch.cuda.synchronize()
for rep in range(2):
    T = {}
    t0 = time.perf_counter()
    def lap(name, t=[t0]):
        now = time.perf_counter(); T[name] = round(now - t[0], 4); t[0] = now
    s = D.Session(0, precond=1); lap("session")
    s.set_mesh(mesh); lap("set_mesh")
    s.upload_ddpe(pinned); lap("upload_ddpe")
    s.ddpe_begin(pr["rec_mode"], pr["delta"], pr["tau_p"], pr["tau_n"], pr["tol"]); lap("ddpe_begin")
    s.ddpe_step(steps); lap("steps")
    s.ddpe_fetch(); lap("fetch")
    st = s.stats()
    T["total"] = round(time.perf_counter() - t0, 4)
    T["lib_t_setup_ms"] = st["t_setup_ms"]; T["lib_t_amg_setup_ms"] = st["t_amg_setup_ms"]
    s.close()
    print(json.dumps(T), flush=True)
