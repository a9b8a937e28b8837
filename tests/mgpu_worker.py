"""Multi-GPU worker (launched by tests/test_multigpu.py under torch.distributed.run): solves the
fixture problems node-partitioned across WORLD_SIZE GPUs and checks them against the golden fields."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ddps_b200 as D  # noqa: E402
from ddps_b200 import _lib  # noqa: E402
from util import KEYS, golden, gpu_mesh, rel_l2  # noqa: E402


def new_comm(rank, world):
    ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (ctypes.c_char * 128)()
        assert _lib.load().ddps_comm_unique_id(buf) == 0
        ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
    dist.broadcast(ident, 0)
    return (bytes(ident.cpu().numpy().tobytes()), rank, world)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # 1. semlin on F1
    mesh = gpu_mesh("F1")
    pr = D.problems.semlin_test(True)
    with D.Session(local, comm=new_comm(rank, world)) as s:
        s.set_mesh(mesh)
        s.setup_semlin(pr["A"], pr["bddata"], pr["f"], pr["g"])
        V, st = s.solve_semlin(pr["tol"])
    z = golden("semlin_F1")
    assert rel_l2(V, z["V"]) <= 1e-8, rel_l2(V, z["V"])
    assert st["outer_iters"] == 5
    # 2. ddpe2 (shockley + Neumann) on F2
    mesh = gpu_mesh("F2")
    pr = D.problems.ddpe2()
    with D.Session(local, comm=new_comm(rank, world)) as s:
        s.set_mesh(mesh)
        s.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
        Vg, ug, vg, st = s.solve_ddpe(pr["rec_mode"], pr["delta"], pr["tau_p"], pr["tau_n"], pr["tol"])
    z = golden("ddpe2_F2")
    for a, k in ((Vg, "V"), (ug, "u"), (vg, "v")):
        assert rel_l2(a, z[k]) <= 1e-8, (k, rel_l2(a, z[k]))
    assert st["outer_iters"] == len(z["control"])
    # 3. structured pn-diode: N-GPU result equals the 1-GPU result (<= 1e-10)
    mesh = D.structured_mesh(96, 48, 0.0, 2.0, 0.0, 1.0)
    pr = D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0)
    with D.Session(local, comm=new_comm(rank, world)) as s:
        s.set_mesh(mesh)
        s.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
        Vn, un, vn, stn = s.solve_ddpe(pr["rec_mode"], pr["delta"], pr["tau_p"], pr["tau_n"], pr["tol"])
    V1, u1, v1, st1 = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], return_stats=True, verbose=False, session=D.Session(local))
    for a, b in ((Vn, V1), (un, u1), (vn, v1)):
        assert rel_l2(a, b) <= 1e-10, rel_l2(a, b)
    assert stn["outer_iters"] == st1["outer_iters"]
    # 4. the DISTRIBUTED multigrid hierarchy (opts.precond = 1) against the single-GPU hierarchy: fields <= 1e-10, equal Gummel
    #    and Newton counts, at most 3 more CG iterations in any solve; forced to build distributed levels on this small mesh
    os.environ["DDPS_AMG_TAIL"] = "1100"
    mesh = D.structured_mesh(192, 96, 0.0, 2.0, 0.0, 1.0)
    with D.Session(local, comm=new_comm(rank, world), precond=1) as s:
        Vn, un, vn, stn = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s, return_stats=True, verbose=False)
    assert stn["amg_levels"] >= 3, stn["amg_levels"]
    with D.Session(local, precond=1) as s:
        V1, u1, v1, st1 = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s, return_stats=True, verbose=False)
    for a, b in ((Vn, V1), (un, u1), (vn, v1)):
        assert rel_l2(a, b) <= 1e-10, rel_l2(a, b)
    assert stn["outer_iters"] == st1["outer_iters"] and list(stn["newton_per_vsolve"]) == list(st1["newton_per_vsolve"])
    assert np.all(np.array(stn["pcg_per_solve"]) <= np.array(st1["pcg_per_solve"]) + 3), (stn["pcg_per_solve"], st1["pcg_per_solve"])
    # 5. per-rank mesh ingest (ddps_mesh_set_local) + owned-node fetch, multigrid path, vs the same single-GPU solve
    part = D.structured_mesh_part(192, 96, 0.0, 2.0, 0.0, 1.0, rank, world)
    with D.Session(local, comm=new_comm(rank, world), precond=1) as s:
        s.set_mesh_local(part)
        s.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
        s.ddpe_begin(pr["rec_mode"], pr["delta"], pr["tau_p"], pr["tau_n"], pr["tol"])
        ctrl = []
        while not ctrl or ctrl[-1] > pr["tol"]:
            ctrl += list(s.ddpe_step(1, stop_at_tol=True))
        Vo, uo, vo = s.ddpe_fetch_owned()
    g = part["own_glob"]
    assert len(ctrl) == st1["outer_iters"]
    for a, b in ((Vo, V1[g]), (uo, u1[g]), (vo, v1[g])):
        assert np.linalg.norm(a - b) <= 1e-10 * np.linalg.norm(b)
    # 6. the same with the NCCL transport instead of the peer-memory kernels
    os.environ["DDPS_P2P"] = "0"
    with D.Session(local, comm=new_comm(rank, world), precond=1) as s:
        Vq, uq, vq, stq = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s, return_stats=True, verbose=False)
    del os.environ["DDPS_P2P"]
    for a, b in ((Vq, V1), (uq, u1), (vq, v1)):
        assert rel_l2(a, b) <= 1e-10, rel_l2(a, b)
    dist.barrier()
    if rank == 0:
        print(f"MGPU_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
