#!/usr/bin/env python
"""Benchmark of the DDPS hot path on B200 (contract: see the task statement / DESIGN.md section 6).

  python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA library through the C ABI)
  python bench.py --impl reference --steps K --warmup W    # the CPU oracle ("port" of the reference; Julia is absent)

A "step" is ONE Gummel iteration (one application of G, src:596-601: Vsolve Newton + two continuity solves) of
solve_ddpe on the synthetic pn-diode of SURVEY.md section 8d config 4: structured 4096x2048-cell triangulation of
[0,2]x[0,1] (16,777,216 triangles), contacts left/right, :shockley.  metric = Gummel iterations per second.

What is timed (round 2: no post-convergence iteration is ever timed):
  value  the K timed steps are the iterations 1..n of WHOLE solves from the zero state to tol (src:649: `while control >
         tol`), repeated until K iterations have run (the last solve is cut at K); device time (CUDA events on the
         library's stream) around the Gummel iterations only, max over ranks; the one-time per-solve setup (base
         stiffness + multigrid hierarchy) is reported separately (`setup_ms_per_solve`), like SURVEY 8d defines the headline.
  e2e    the drop-in call a user makes, `D.solve_ddpe(mesh, closures..., tol)`, cold, wall clock: context creation, mesh H2D
         from pinned host arrays, pattern build, coefficient sampling, hierarchy setup, Gummel to tol, D2H of V,u,v.
Inputs are > L2 (the matrix alone is 705 MB), so no L2 flush is needed between iterations.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"s16": (4096, 2048), "s1": (1024, 512), "s4": (2048, 1024), "s64": (8192, 4096), "s256": (16384, 8192),
             "s025": (512, 256), "tiny": (128, 64)}
CPU_SIZES = [(128, 64), (256, 128), (512, 256)]      # what the CPU oracle solves to tol in seconds..a minute
METRIC = "gummel_iters_per_s"
KEYS = ("rec_mode", "Vbddata", "ubddata", "vbddata", "lam", "delta", "tau_p", "tau_n", "mu_p", "mu_n", "C", "tol")


def problem():
    import ddps_b200 as D
    return D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)


def workload_name(nx, ny):
    return (f"solve_ddpe pn-diode (C0=1, U=0.3, :shockley, tol 1e-9), structured {nx}x{ny} cells "
            f"({2 * nx * ny} triangles, {(nx + 1) * (ny + 1)} nodes), one Gummel iteration per step, whole solves from the zero state to tol")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- CPU oracle legs
def oracle_gummel(nx, ny, warmup, steps):
    """The CPU oracle (numpy + scipy SuperLU, shim S-a) on an nx x ny-cell mesh: `warmup` untimed Gummel iterations, then
    `steps` timed ones taken from whole solves to tol (the same window the GPU arm times).  Returns seconds per iteration
    MEASURED at this size plus the iterations one solve needs."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ddps_oracle as O
    pr = problem()
    m = O.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0)
    _, _, gD, n = O.aquire_boundary(m, pr["Vbddata"])
    MV, bdMV, ar = O.MV_assemble(m, n, gD, pr["lam"])[:3]
    nq = m.nodes.shape[1]

    def run(limit, timed):
        """Gummel iterations of whole solves until `limit` iterations have run; returns (seconds, iterations per solve)."""
        done, sec, per_solve = 0, 0.0, []
        while done < limit:
            u = np.zeros(nq); v = np.zeros(nq); V = np.zeros(nq)
            control, k = 1.0, 0
            while control > pr["tol"] and done < limit:
                t0 = time.perf_counter()
                u0, v0, V0 = u, v, V
                u, v, V = O.G(m, pr["rec_mode"], u0, v0, pr["delta"], pr["tau_p"], pr["tau_n"], pr["mu_p"], pr["mu_n"], gD, n, MV,
                              bdMV, ar, pr["ubddata"], pr["vbddata"], pr["C"], pr["tol"])
                control = max(np.max(np.abs(u - u0)), np.max(np.abs(v - v0)), np.max(np.abs(V - V0)))
                sec += time.perf_counter() - t0
                k += 1
                done += 1
            if not control > pr["tol"]:
                per_solve.append(k)
        return sec, per_solve

    if warmup > 0:
        run(warmup, False)
    sec, per_solve = run(steps, True)
    return dict(cells=[nx, ny], triangles=2 * nx * ny, steps=steps, sec_per_iter=sec / steps, iters_to_tol=per_solve[0] if per_solve else None)


def cpu_scaling_fit(samples):
    """Least-squares exponent p of sec_per_iter ~ triangles^p over the measured sizes."""
    x = np.log([s["triangles"] for s in samples])
    y = np.log([s["sec_per_iter"] for s in samples])
    p, c = np.polyfit(x, y, 1)
    return float(p), float(c)


def run_reference(args, nx, ny):
    """Reference arm: the reference is Julia (absent from the image: no `julia`, nothing to compile), so this arm times the
    CPU oracle -- the restatement of the same functions -- on the box's host cores.  A full Gummel iteration at the
    workload's size is hours of sparse LU, so each step is a bounded sample of the workload: ONE Gummel iteration on the
    512x256-cell mesh (the same diode, the same whole-solves-to-tol window).  `value` and `ms_per_step` are what was
    MEASURED at that size (no scaling); `config.workload` says so; the extrapolations to the workload are separate keys."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    snx, sny = CPU_SIZES[-1] if 2 * nx * ny >= 2 * CPU_SIZES[-1][0] * CPU_SIZES[-1][1] else (nx, ny)
    t0 = time.perf_counter()
    small = [oracle_gummel(a, b, 0, 10) for a, b in CPU_SIZES[:-1] if 2 * a * b < 2 * snx * sny]
    main = oracle_gummel(snx, sny, args.warmup, args.steps)
    rate = 1.0 / main["sec_per_iter"]
    work_tri = 2 * nx * ny
    ext = None
    if len(small) >= 2:
        p, c = cpu_scaling_fit(small + [main])
        ext = {"exponent_fitted": p, "sizes": [s["triangles"] for s in small + [main]],
               "sec_per_iter": [s["sec_per_iter"] for s in small + [main]],
               "rate_at_workload_linear": rate * main["triangles"] / work_tri,
               "rate_at_workload_fitted": 1.0 / float(np.exp(c) * work_tri ** p)}
    sample = (f"{args.warmup}+{args.steps} Gummel iterations (whole solves to tol, {main['iters_to_tol']} iterations each) of the CPU oracle "
              f"(numpy + scipy SuperLU, shim S-a, 1 thread) on the {snx}x{sny}-cell diode ({main['triangles']} triangles = "
              f"1/{work_tri // main['triangles']} of our arm's workload); value is the rate MEASURED at that size")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "1/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * main["sec_per_iter"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(snx, sny),
                       "bounded_sample_of": workload_name(nx, ny),
                       "note": "the CPU cannot run the 16.8M-triangle workload inside a bench run (hours of sparse LU per Gummel iteration): "
                               "this arm runs the same diode at the largest size it solves in minutes; our arm's line carries the measured "
                               "same-size pair (`same_size_pair`)"},
            "cpu_baseline": {"value": rate, "unit": "1/s", "cores": 1, "kind": "port", "sample": sample,
                             "host_cores_visible": os.cpu_count(), "extrapolation_to_our_workload": ext},
            "e2e": {"value": rate, "unit": "1/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_seconds": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- our arm
def new_comm(torch, dist, rank, world):
    from ddps_b200 import _lib
    ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (ctypes.c_char * 128)()
        assert _lib.load().ddps_comm_unique_id(buf) == 0
        ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
    dist.broadcast(ident, 0)
    return (bytes(ident.cpu().numpy().tobytes()), rank, world)


def timed_solves(s, pr, limit, barrier, allmax):
    """Gummel iterations 1..n of whole solves to tol until `limit` iterations have run.  Device milliseconds of the
    iterations (library events, max over ranks), the setup milliseconds of the solves, launches, CG iterations, controls."""
    done, ms, setup_ms, launches, pcg, ctrl_all, per_solve = 0, 0.0, 0.0, 0, 0, [], []
    while done < limit:
        barrier()
        t0 = time.perf_counter()
        s.ddpe_begin(pr["rec_mode"], pr["delta"], pr["tau_p"], pr["tau_n"], pr["tol"])     # base stiffness + hierarchy (per solve)
        barrier()
        setup_ms += 1000.0 * (time.perf_counter() - t0)
        st0 = s.stats()
        ctrl = s.ddpe_step(limit - done, stop_at_tol=True)
        barrier()
        st1 = s.stats()
        ms += allmax(st1["t_solve_ms"] - st0["t_solve_ms"])
        launches += st1["kernel_launches"] - st0["kernel_launches"]
        pcg += st1["pcg_total"] - st0["pcg_total"]
        done += len(ctrl)
        ctrl_all += [float(c) for c in ctrl]
        if len(ctrl) and not ctrl[-1] > pr["tol"]:
            per_solve.append(len(ctrl))
    return dict(ms=ms, setup_ms=setup_ms, launches=int(launches), pcg=int(pcg), control=ctrl_all, per_solve=per_solve, nsolves=max(1, len(per_solve)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="s16", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel roofline timings")
    ap.add_argument("--precond", default="amg", choices=["amg", "jacobi"],
                    help="preconditioner of the hand-written CG: amg = smoothed-aggregation multigrid V-cycle (default), "
                         "jacobi = the diagonal preconditioner the north star names")
    ap.add_argument("--digest", default="", help="after the timed solves run one more whole solve and write every 997th node's V,u,v "
                                                 "(global ids) of this rank to <digest>_n<N>_r<rank>.npz: cross-checks between runs")
    ap.add_argument("--no-lag", action="store_true", help="opts.amg_no_lag=1 (saves the per-kind operator copies: memory)")
    ap.add_argument("--pcg-cap", type=int, default=0,
                    help="PROFILING ONLY: cap every PCG solve at this many iterations (keeps an ncu launch list short); "
                         "the printed value is then not a benchmark number")
    args = ap.parse_args()
    nx, ny = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, nx, ny)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import ddps_b200 as D
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    precond = args.precond
    sopts = dict(precond=1) if precond == "amg" else {}
    if args.no_lag:
        sopts["amg_no_lag"] = 1
    pr = problem()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    if world == 1:
        mesh0 = D.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0)
        # the caller's arrays in pinned host memory, in the ABI's layout (= the reference's Julia arrays: nodes 2 x nq and
        # elements 5 x nme column-major); the Mesh handed to the library is a zero-copy view of them
        nodes, elems = pin(mesh0.nodes.T), pin(mesh0.elements.T)
        mesh = D.Mesh(nodes.T, mesh0.edges, elems.T)
        nq = mesh.nq
        del mesh0
        part = None
        h2d_mesh = nodes.nbytes + elems.nbytes
    else:
        # multi-GPU: per-rank mesh ingest -- every rank builds and uploads ONLY its strip of the mesh (ddps_mesh_set_local);
        # nobody holds the global arrays (13 GB per process at 268M triangles)
        part = D.structured_mesh_part(nx, ny, 0.0, 2.0, 0.0, 1.0, rank, world)
        for k in ("nodes", "elements", "node_glob"):
            part[k] = pin(part[k])
        part["elem_glob"] = None     # coefficients are sampled on the device: no global triangle ids needed
        mesh = None
        nq = part["nq_global"]
        h2d_mesh = part["nodes"].nbytes + part["elements"].nbytes + part["node_glob"].nbytes

    def fresh_session():
        return D.Session(local, comm=new_comm(torch, dist, rank, world) if dist is not None else None, **sopts)

    def load(s):
        if part is None:
            s.set_mesh(mesh)
        else:
            s.set_mesh_local(part)
        s.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])

    # ---- e2e FIRST (cold): D.solve_ddpe(mesh, closures..., tol) -> (V,u,v) -----------------------------------------
    e2e = None
    if not args.no_e2e and args.pcg_cap == 0:
        # one-time PROCESS initialisation is not part of the measurement: a small solve creates the CUDA context, loads the
        # library's kernels (lazy module loading: 131,841 rows are enough for the large-input variants of the setup's sort / scan
        # kernels and for the device-side coarsening, > 100,000 rows, to be loaded) and creates the memory pool
        with D.Session(local, **sopts) as sw:
            D.solve_ddpe(D.structured_mesh(512, 256, 0.0, 2.0, 0.0, 1.0), *[pr[k] for k in KEYS], session=sw, verbose=False)
        barrier()
        t0 = time.perf_counter()
        s2 = fresh_session()
        if part is None:
            V, u, v, st = D.solve_ddpe(mesh, *[pr[k] for k in KEYS], session=s2, return_stats=True, verbose=False)
            iters = int(st["outer_iters"])
        else:
            # the same call sequence solve_ddpe makes, on this rank's part of the mesh; every rank fetches the nodes it owns
            load(s2)
            s2.ddpe_begin(pr["rec_mode"], pr["delta"], pr["tau_p"], pr["tau_n"], pr["tol"])
            iters = len(s2.ddpe_step(1000, stop_at_tol=True))
            V, u, v = s2.ddpe_fetch_owned()
            st = s2.stats()
        barrier()
        t_e2e = allmax(time.perf_counter() - t0)
        nd = len(D.dirichlet_nodes(mesh if part is None else part["boundary"], pr["Vbddata"])[0])
        h2d = h2d_mesh + 3 * 16 * nd      # mesh + three Dirichlet (node,value) lists; coefficients sampled on the device
        d2h = 3 * 8 * (nq if part is None else part["n_own"])
        e2e = {"value": iters / t_e2e, "unit": "1/s", "h2d_bytes_per_step": h2d / iters, "d2h_bytes_per_step": d2h / iters,
               "seconds": t_e2e, "gummel_iters_to_tol": iters,
               "what": ("cold D.solve_ddpe(mesh, closures..., tol) -> (V,u,v): " if part is None else
                        "cold solve_ddpe call sequence on per-rank mesh parts (bytes: this rank's), every rank fetches its owned nodes: ")
                       + "context creation, mesh H2D from pinned host arrays, pattern build, boundary preprocessing on the host, "
                         "coefficient closures (registered spatial functors) sampled on the device, base stiffness + multigrid "
                         "hierarchy setup, Gummel to tol, D2H of V,u,v",
               "setup_seconds": st["t_setup_ms"] / 1000.0 + st["t_amg_setup_ms"] / 1000.0, "final_control": float(st["control"])}
        s2.close()
        del V, u, v

    # ---- resident: whole solves to tol, device-timed ------------------------------------------------------------------
    s = fresh_session()
    if args.pcg_cap > 0:
        s.set_opts(lin_max_it=args.pcg_cap, lin_cap_ok=1, max_newton=3)
    load(s)
    timed_solves(s, pr, args.warmup, barrier, allmax)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    R = timed_solves(s, pr, args.steps, barrier, allmax)
    clocks = sampler.stop() if rank == 0 else None
    value = args.steps / (R["ms"] / 1000.0)
    levels = s.amg_levels() if precond == "amg" else []
    amg_built = s.stats()["amg_levels"]
    free_b, total_b = torch.cuda.mem_get_info()
    mem_used_gb = allmax((total_b - free_b) / 2 ** 30)
    if args.digest:
        s.ddpe_begin(pr["rec_mode"], pr["delta"], pr["tau_p"], pr["tau_n"], pr["tol"])
        nit = len(s.ddpe_step(1000, stop_at_tol=True))
        if part is None:
            Vd, ud, vd = s.ddpe_fetch()
            gid = np.arange(nq, dtype=np.int64)
        else:
            Vd, ud, vd = s.ddpe_fetch_owned()
            gid = part["own_glob"]
        pick = (gid % 997) == 0
        np.savez(f"{args.digest}_n{world}_r{rank}.npz", gid=gid[pick], V=Vd[pick], u=ud[pick], v=vd[pick], iters=nit,
                 sums=np.array([Vd.sum(), ud.sum(), vd.sum(), (Vd * Vd).sum(), (ud * ud).sum(), (vd * vd).sum()]))
        del Vd, ud, vd

    # ---- roofline: every kernel of the benchmarked iteration, timed live in this process (CUDA events on the library
    #      stream, ddps_time_kernel); share = launches per step x ms per launch / ms per step -------------------------------
    peak, peak_src = measured_peak()
    roofline = None
    if not args.no_kernels and world == 1:
        traffic = {}
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get(args.workload, {})
        except Exception:
            pass
        ms_step = R["ms"] / args.steps
        its_step = R["pcg"] / args.steps
        solves_step = (sum(s.history(1)) + 2 * len(s.history(1))) / max(1, len(s.history(1)))   # Newton systems + 2 continuity systems per Gummel iteration

        def entry(kid, name, per_step, warm=3, reps=30, tkey=None):
            ms, by = s.time_kernel(kid, warm, reps)
            return {"kernel": name, "ms_per_launch": ms, "algo_bytes_per_launch": by, "achieved": by / ms / 1e6 if by else None,
                    "frac": by / ms / 1e6 / peak if by else None, "launches_per_step": per_step, "share_of_step": per_step * ms / ms_step,
                    "traffic": traffic.get(tkey) if tkey else None}

        kern = {"spmv_dot": entry(5, "k_spmv<true,false> (SELL-32 SpMV + fused p.Ap, fp64 values)", its_step, tkey="k_spmv_dot"),
                "assembly_uv": entry(1, "k_prep_uv + k_asm_stiff<EXPV,MASS,SRH> (continuity system, 72 B/triangle)", 2, tkey="k_assemble_uv"),
                "assembly_newton": entry(3, "k_prep_poisson + k_asm_newton<DDP> (Newton system: A = MV - J0, b, F = MV*V - b; 134 B/triangle)", solves_step - 2 + 1,
                                         tkey="k_assemble_newton")}
        if precond == "amg":
            kern.update({
                "amg_resid_spmv": entry(9, "k_amg_spmv<1,false,float> finest level (t = b - A x, fp32 matrix copy)", its_step, tkey="k_amg_spmv_1"),
                "amg_smooth_spmv": entry(10, "k_amg_spmv<2,true,float> finest level (Jacobi sweep + r.z)", its_step, tkey="k_amg_spmv_2"),
                "apcg_update": entry(11, "k_apcg_update (x, r, first sweep)", its_step),
                "restrict_0": entry(12, "k_amg_restrict level 0 -> 1", its_step),
                "prolong_0": entry(13, "k_amg_prolong level 1 -> 0", its_step),
                "cg_iteration": entry(6, "one multigrid-preconditioned CG iteration (all levels, all launches)", its_step, 2, 20),
                "numeric_phase": entry(7, "per-operator numeric phase (Galerkin products, fp32 copies, coarsest inverse); skipped when the lagged coarse operators are reused", 0, 2, 10)})
            kern["vcycle_ms"] = s.time_kernel(8, 2, 20)[0]
        else:
            kern["pcg_iteration"] = entry(2, "one Jacobi-PCG iteration (3 kernels)", its_step)
        cands = [k for k, e in kern.items() if isinstance(e, dict) and e.get("frac") and k not in ("cg_iteration", "pcg_iteration", "numeric_phase")]
        dom = max(cands, key=lambda k: kern[k]["share_of_step"])
        d = kern[dom]
        roofline = {"bound": "hbm", "kernel": d["kernel"], "achieved": d["achieved"], "peak": peak, "unit": "GB/s", "frac": d["frac"],
                    "traffic": d["traffic"], "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                    "peak_source": peak_src, "algo_bytes_per_launch": d["algo_bytes_per_launch"], "ms_per_launch": d["ms_per_launch"],
                    "share_of_step": d["share_of_step"], "dominant": dom,
                    "why": "the single kernel with the largest share of the timed step; every other kernel of the step is listed in `kernels`",
                    "kernels": kern, "cg_iterations_per_step": its_step, "ms_per_step": ms_step}

    # ---- same-size pair: this GPU arm at the largest size the CPU oracle solves, next to the oracle measured here ----------
    cpu, pair = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.pcg_cap == 0:
        snx, sny = CPU_SIZES[-1]
        ms_ = D.structured_mesh(snx, sny, 0.0, 2.0, 0.0, 1.0)
        with D.Session(local, **sopts) as sp:
            sp.set_mesh(ms_)
            sp.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
            timed_solves(sp, pr, 3, barrier, allmax)
            Rp = timed_solves(sp, pr, 10, barrier, allmax)
        t0 = time.perf_counter()
        samples = [oracle_gummel(a, b, 0, 10) for a, b in CPU_SIZES[:-1]]
        big = oracle_gummel(snx, sny, 0, 10)
        p, c = cpu_scaling_fit(samples + [big])
        work_tri = 2 * nx * ny
        gpu_rate = 10 / (Rp["ms"] / 1000.0)
        pair = {"workload": workload_name(snx, sny), "gpu_value": gpu_rate, "cpu_value": 1.0 / big["sec_per_iter"], "unit": "1/s",
                "ratio": gpu_rate * big["sec_per_iter"], "same_config": True,
                "what": "both sides MEASURED in this run on the same mesh and the same window (10 Gummel iterations of whole solves to tol): "
                        "the resident GPU path vs the CPU oracle (1 thread: scipy SuperLU and numpy are single-threaded)"}
        cpu = {"value": 1.0 / big["sec_per_iter"], "unit": "1/s", "cores": 1, "kind": "port",
               "sample": (f"10 Gummel iterations (whole solves to tol) of the CPU oracle (numpy + scipy SuperLU) on the {snx}x{sny}-cell diode "
                          f"({big['triangles']} triangles = 1/{work_tri // big['triangles']} of the workload): value is the rate MEASURED at that size"),
               "host_cores_visible": os.cpu_count(),
               "extrapolation_to_workload": {"exponent_fitted": p, "sizes": [x["triangles"] for x in samples + [big]],
                                             "sec_per_iter": [x["sec_per_iter"] for x in samples + [big]],
                                             "rate_linear": (1.0 / big["sec_per_iter"]) * big["triangles"] / work_tri,
                                             "rate_fitted": 1.0 / float(np.exp(c) * work_tri ** p)},
               "wall_seconds": time.perf_counter() - t0}
    s.close()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "1/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": R["ms"] / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(nx, ny),
                           "l2": "inputs larger than L2 (CSR values+columns 705 MB at s16); no flush needed",
                           "parallelism": (f"node-partitioned x{world} (vertical strips, one process per GPU, per-rank mesh ingest; distributed "
                                           f"multigrid hierarchy: {amg_built} levels, rows owned by rank 0 per level {levels})") if world > 1 else "single GPU",
                           "solver": ("hand-written CG preconditioned with a smoothed-aggregation multigrid V(1,1) cycle "
                                      "(opts.precond=1; SURVEY 8f rank 1), lin_rtol 1e-13, Newton corrections to 0.1*tol"
                                      if precond == "amg" else "Jacobi-PCG, lin_rtol 1e-13, Newton corrections to 0.1*tol"),
                           "preconditioner": precond,
                           "timed_window": "Gummel iterations 1..n of whole solves from the zero state to tol (no post-convergence iteration)",
                           "precision": ("CG operator, vectors, dots and all arithmetic f64; the multigrid cycle (a preconditioner only) "
                                         "stores its transfer operators and a copy of the large levels' matrix values in fp32; coarse "
                                         "operators lagged per operator kind (opts.amg_no_lag=0); parity tests: fields <= 1e-8 vs the oracle"
                                         if precond == "amg" else "f64 throughout")},
                "clocks": clocks, "e2e": e2e, "gpu_launches": R["launches"], "pcg_iterations": R["pcg"],
                "gummel_iters_to_tol": R["per_solve"][0] if R["per_solve"] else None, "solves_timed": R["nsolves"],
                "setup_ms_per_solve": R["setup_ms"] / R["nsolves"], "amg_levels": levels, "gpu_mem_used_gb_max_over_ranks": mem_used_gb,
                "control": R["control"], "roofline": roofline, "cpu_baseline": cpu, "same_size_pair": pair}
        if args.pcg_cap > 0:
            line["invalid"] = f"profiling run: PCG capped at {args.pcg_cap} iterations per solve"
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
