#!/usr/bin/env python
"""Benchmark of the DDPS hot path on B200 (contract: see the task statement / DESIGN.md section 6).

  python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA library through the C ABI)
  python bench.py --impl reference --steps K --warmup W    # the CPU oracle ("port" of the reference)

A "step" is ONE Gummel iteration (one application of G, src:596-601: Vsolve Newton + two continuity
solves) of solve_ddpe on the synthetic pn-diode of SURVEY.md section 8d config 4: structured
4096x2048-cell triangulation of [0,2]x[0,1] (16,777,216 triangles), contacts left/right, :shockley.
metric = Gummel iterations per second.  Inputs are > L2 (the matrix alone is 705 MB), so no L2 flush
is needed between iterations.
"""
import argparse
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
             "tiny": (128, 64)}
METRIC = "gummel_iters_per_s"


def problem():
    import ddps_b200 as D
    return D.problems.pn_junction([2, 4], [1, 2, 3, 4], x_junction=1.0, C0=1.0, U=0.3)


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


def cpu_oracle_rate(nx, ny, iters, work_tri):
    """Time `iters` Gummel iterations of the CPU oracle on an nx x ny sample and scale the rate to the
    workload's triangle count LINEARLY (optimistic for the CPU: its sparse factorisations scale
    super-linearly, measured exponent ~1.25-1.45)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ddps_oracle as O
    m = O.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0)
    t0 = time.perf_counter()
    sec = O.timed_gummel_iterations(m, problem(), iters)
    tri = m.elements.shape[1]
    return dict(sec_per_iter_sample=sec, sample_triangles=tri, wall=time.perf_counter() - t0,
                rate_sample=1.0 / sec, rate_scaled=(1.0 / sec) * tri / work_tri)


def run_reference(args, nx, ny):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    work_tri = 2 * nx * ny
    snx, sny = (512, 256) if work_tri >= 2 * 512 * 256 else (nx, ny)
    for _ in range(0):
        pass
    t0 = time.perf_counter()
    r = cpu_oracle_rate(snx, sny, max(args.steps + min(args.warmup, 1), 1), work_tri)
    sample = (f"{args.steps}+{min(args.warmup, 1)} Gummel iterations of the CPU oracle (numpy + scipy SuperLU, shim S-a) on a "
              f"{snx}x{sny}-cell mesh ({r['sample_triangles']} triangles = 1/{work_tri // r['sample_triangles']} of the workload); "
              f"rate scaled linearly in triangles to the workload (optimistic for the CPU)")
    line = {"impl": "reference", "metric": METRIC, "value": r["rate_scaled"], "unit": "1/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / r["rate_scaled"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"solve_ddpe pn-diode, structured {nx}x{ny} cells ({work_tri} triangles), one Gummel iteration per step"},
            "cpu_baseline": {"value": r["rate_scaled"], "unit": "1/s", "cores": 1, "kind": "port", "sample": sample,
                             "sec_per_iter_on_sample": r["sec_per_iter_sample"]},
            "e2e": {"value": r["rate_scaled"], "unit": "1/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="s16", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--precond", default="auto", choices=["auto", "amg", "jacobi"],
                    help="preconditioner of the hand-written CG: amg = smoothed-aggregation multigrid V-cycle (single GPU), "
                         "jacobi = the diagonal preconditioner; auto = amg on one GPU, jacobi with a communicator")
    ap.add_argument("--pcg-cap", type=int, default=0,
                    help="PROFILING ONLY: cap every PCG solve at this many iterations (keeps an ncu launch list short); "
                         "the printed value is then not a benchmark number")
    args = ap.parse_args()
    nx, ny = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, nx, ny)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import ddps_b200 as D
    comm = None
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from ddps_b200 import _lib
        ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (__import__("ctypes").c_char * 128)()
            assert _lib.load().ddps_comm_unique_id(buf) == 0
            ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(ident, 0)
        comm = (bytes(ident.cpu().numpy().tobytes()), rank, world)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    precond = args.precond if args.precond != "auto" else ("amg" if world == 1 else "jacobi")
    if precond == "amg" and world > 1:
        precond = "jacobi"   # the multigrid hierarchy is single-GPU in this round
    sopts = dict(precond=1) if precond == "amg" else {}
    pr = problem()
    mesh = D.structured_mesh(nx, ny, 0.0, 2.0, 0.0, 1.0)
    smp = D.api.sample_ddpe(mesh, pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
    def measure_e2e():
        # end to end through the C ABI with host buffers: mesh + coefficient uploads, setup, K steps, D2H of V,u,v
        e2e = None
        if not args.no_e2e:
            # (measured FIRST, in the fresh process: this is the cold start a user sees; a session created after ~10 GB of
            # cudaFree runs into a slow and erratic legacy allocation path, which would be timed as "setup")
            # host buffers in the ABI's layout (= the memory layout of the reference's Julia arrays: nodes 2 x nq and elements
            # 5 x nme column-major), pinned; the Mesh handed to the library is a zero-copy view of them
            nodes = torch.from_numpy(np.ascontiguousarray(mesh.nodes.T)).pin_memory().numpy()
            elems = torch.from_numpy(np.ascontiguousarray(mesh.elements.T)).pin_memory().numpy()
            mesh_p = D.Mesh(nodes.T, mesh.edges, elems.T)
            pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in smp.items()}
            h2d = nodes.nbytes + elems.nbytes + sum(pinned[k].nbytes for k in ("lambda2", "mu_p", "mu_n", "C_mid")) + 3 * 16 * len(smp["n"])
            d2h = 3 * 8 * mesh.nq
            outs = tuple(torch.empty(mesh.nq, dtype=torch.float64).pin_memory().numpy() for _ in range(3))   # caller-owned outputs
            if comm is None:
                # one-time process initialisation is not part of the measurement: a tiny solve loads the library's kernels
                # (lazy module loading) and creates the memory pool before the timed cold start of the real workload
                mt = D.structured_mesh(64, 32, 0.0, 2.0, 0.0, 1.0)
                with D.Session(local, **sopts) as sw:
                    sw.set_mesh(mt)
                    sw.setup_ddpe(pr["Vbddata"], pr["ubddata"], pr["vbddata"], pr["lam"], pr["mu_p"], pr["mu_n"], pr["C"])
                    sw.ddpe_begin(pr["rec_mode"], pr["delta"], pr["tau_p"], pr["tau_n"], pr["tol"])
                    sw.ddpe_step(1)
                    sw.ddpe_fetch()
            barrier()
            t0 = time.perf_counter()
            s2 = D.Session(local, comm=comm, **sopts) if comm is None else None
            if s2 is None:   # multi-GPU: a communicator cannot be re-created from the same id; reuse a fresh one
                ident2 = torch.zeros(128, dtype=torch.uint8, device="cuda")
                if rank == 0:
                    buf = (__import__("ctypes").c_char * 128)()
                    _lib.load().ddps_comm_unique_id(buf)
                    ident2 = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
                dist.broadcast(ident2, 0)
                s2 = D.Session(local, comm=(bytes(ident2.cpu().numpy().tobytes()), rank, world))
            s2.set_mesh(mesh_p)
            s2.upload_ddpe(pinned)
            s2.ddpe_begin(pr["rec_mode"], pr["delta"], pr["tau_p"], pr["tau_n"], pr["tol"])
            s2.ddpe_step(args.steps)
            s2.ddpe_fetch(out=outs)
            barrier()
            t_e2e = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                t_e2e = float(t.item())
            e2e = {"value": args.steps / t_e2e, "unit": "1/s", "h2d_bytes_per_step": h2d / args.steps,
                   "d2h_bytes_per_step": d2h / args.steps, "seconds": t_e2e,
                   "note": "cold start: mesh+coefficient H2D from pinned host buffers, pattern build"
                           + (", multigrid hierarchy setup (device, greedy aggregation on the host)" if precond == "amg" else "") + ", first K Gummel iterations, D2H of V,u,v",
                   "setup_seconds": s2.stats()["t_setup_ms"] / 1000.0 + s2.stats()["t_amg_setup_ms"] / 1000.0}
            s2.close()
        return e2e

    # single GPU: the cold start is measured FIRST, in the fresh process; with a communicator the validated order is kept
    # (resident session first, then the end-to-end session with a second communicator)
    e2e = measure_e2e() if comm is None else None

    s = D.Session(local, comm=comm, **sopts)
    if args.pcg_cap > 0:
        s.set_opts(lin_max_it=args.pcg_cap, lin_cap_ok=1, max_newton=3)
    s.set_mesh(mesh)
    s.upload_ddpe(smp)
    s.ddpe_begin(pr["rec_mode"], pr["delta"], pr["tau_p"], pr["tau_n"], pr["tol"])
    if args.warmup > 0:
        s.ddpe_step(args.warmup)
    st0 = s.stats()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctrl = s.ddpe_step(args.steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    st1 = s.stats()
    dev_ms = st1["t_solve_ms"] - st0["t_solve_ms"]
    if dist is not None:
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    value = args.steps / (dev_ms / 1000.0)
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    pcg_its = st1["pcg_total"] - st0["pcg_total"]

    # roofline of the dominant kernel: the PCG's SELL-32 SpMV with fused dot (52% of a PCG iteration)
    peak, peak_src = measured_peak()
    ms_spmv, by_spmv = s.time_kernel(5, 5, 50)
    ms_pcg, by_pcg = s.time_kernel(2, 5, 50)
    ms_asm, by_asm = s.time_kernel(1, 3, 20)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(args.workload, {}).get("k_spmv_dot")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_spmv<true> (SELL-32 SpMV + fused p.Ap)", "achieved": by_spmv / ms_spmv / 1e6,
                "peak": peak, "unit": "GB/s", "frac": by_spmv / ms_spmv / 1e6 / peak, "traffic": traffic,
                "peak_source": peak_src, "algo_bytes_per_launch": by_spmv, "ms_per_launch": ms_spmv,
                "pcg_iteration": {"achieved": by_pcg / ms_pcg / 1e6, "frac": by_pcg / ms_pcg / 1e6 / peak, "ms": ms_pcg,
                                  "what": "one Jacobi-PCG iteration (3 kernels)"},
                "assembly_uv": {"achieved": by_asm / ms_asm / 1e6, "frac": by_asm / ms_asm / 1e6 / peak, "ms": ms_asm}}
    if precond == "amg":
        # the multigrid-preconditioned iteration: fine-level SpMV x3 (this kernel and its two fused-epilogue variants) +
        # transfers + coarse levels; the per-operator numeric phase (Galerkin products + coarsest inverse)
        ms_it, _ = s.time_kernel(6, 2, 20)
        ms_vc, _ = s.time_kernel(8, 2, 20)
        ms_num, _ = s.time_kernel(7, 2, 10)
        roofline["amg"] = {"levels": s.amg_levels(), "ms_per_cg_iteration": ms_it, "ms_per_vcycle": ms_vc,
                           "ms_numeric_phase_per_operator": ms_num,
                           "fine_spmv_share_of_iteration": 3 * ms_spmv / ms_it,
                           "t_hierarchy_setup_ms": st1["t_amg_setup_ms"]}

    if comm is not None:
        e2e = measure_e2e()
    s.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        work_tri = 2 * nx * ny
        snx, sny = (512, 256) if work_tri >= 2 * 512 * 256 else (nx, ny)
        r = cpu_oracle_rate(snx, sny, 2, work_tri)
        cpu = {"value": r["rate_scaled"], "unit": "1/s", "cores": 1, "kind": "port",
               "sample": (f"2 Gummel iterations of the CPU oracle (numpy + scipy SuperLU) on {snx}x{sny} cells "
                          f"({r['sample_triangles']} triangles); {r['sec_per_iter_sample']:.2f} s/iteration on the sample, rate scaled "
                          f"linearly in triangles to the workload (optimistic for the CPU)"),
               "host_cores_visible": os.cpu_count()}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "1/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"solve_ddpe pn-diode (C0=1, U=0.3, :shockley), structured {nx}x{ny} cells "
                                       f"({2 * nx * ny} triangles, {mesh.nq} nodes), one Gummel iteration per step",
                           "l2": "inputs larger than L2 (CSR values+columns 705 MB at s16); no flush needed",
                           "parallelism": f"node-partitioned x{world}" if world > 1 else "single GPU",
                           "solver": ("hand-written CG preconditioned with a smoothed-aggregation multigrid V(1,1) cycle "
                                      "(opts.precond=1; SURVEY 8f rank 1), lin_rtol 1e-13, Newton corrections to 0.1*tol"
                                      if precond == "amg" else "Jacobi-PCG, lin_rtol 1e-13, Newton corrections to 0.1*tol"),
                           "preconditioner": precond,
                           "precision": ("CG operator, vectors, dots and all arithmetic f64; the multigrid cycle (a preconditioner only) "
                                         "stores its transfer operators and a copy of the large levels' matrix values in fp32; coarse "
                                         "operators lagged per operator kind (opts.amg_no_lag=0); parity tests: fields <= 1e-8 vs the oracle"
                                         if precond == "amg" else "f64 throughout")},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "pcg_iterations": int(pcg_its),
                "control": [float(c) for c in ctrl], "roofline": roofline, "cpu_baseline": cpu}
        if world > 1:
            line["note"] = ("node-partitioned Jacobi-PCG (the multigrid hierarchy is single-GPU in this round, DESIGN.md section 9): on this "
                            "workload ONE GPU with the multigrid-preconditioned CG delivers ~14.6 Gummel it/s (profiles/r01_bench_s16_n1_amg.json)")
        if args.pcg_cap > 0:
            line["invalid"] = f"profiling run: PCG capped at {args.pcg_cap} iterations per solve"
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
