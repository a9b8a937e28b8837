#!/usr/bin/env python
"""Blueprint of the DISTRIBUTED multigrid hierarchy, validated numerically on the CPU.

P ranks are simulated in one process.  Every rank only ever touches (a) the rows it owns, (b) values of its ghost
columns that arrived in an explicit message, and (c) small replicated data; every message is routed through `Net`, which
counts them, so the script is also the list of exchanges the CUDA implementation needs:

  setup, per distributed level          1. aggregate ids of the ghost fine nodes           (halo of one int per node)
                                        2. prolongator rows of the ghost fine nodes        (variable-length rows, once)
                                        3. pattern of A*P for the ghost fine rows my restriction rows touch (once)
  per operator, per distributed level   4. values of A*P for those ghost fine rows         (fixed pattern -> a halo of values)
  per cycle, per distributed level      5. halo(x) before the residual, 6. halo(t) before the restriction,
                                        7. halo(xc) before the prolongation, 8. halo(x) before the smoothing sweep
  tail (levels below --tail rows)       9. allgather of the coarse operator (per operator) and of its right-hand side (per
                                           cycle); the tail of the hierarchy is replicated and solved redundantly.

Checks: (i) the distributed Galerkin operator equals P^T A P of the assembled global prolongator, (ii) one distributed
V-cycle equals the global V-cycle on the same hierarchy to rounding, (iii) CG needs the same number of iterations, and only
0-2 more than with the single-GPU (unconstrained) aggregation.  Test infrastructure, not product.

  python tools/amg_dist_prototype.py 256 128 4 [--tail 3000]
"""
import os
import sys
from collections import defaultdict

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import amg_prototype as PR  # noqa: E402
import amg_ref  # noqa: E402
import ddps_b200 as D  # noqa: E402

OMEGA = 2.0 / 3.0


class Net:
    """Message accounting: kind -> [number of messages, payload items]."""

    def __init__(self):
        self.log = defaultdict(lambda: [0, 0])

    def send(self, kind, src, dst, payload_items):
        if src != dst:
            self.log[kind][0] += 1
            self.log[kind][1] += int(payload_items)


def f32(a):
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)


def local_aggregates(Abase_rows, rows, fixed):
    """Greedy aggregation of the owned rows; couplings to nodes owned by other ranks are invisible to it (the library's
    own aggregation on the owned diagonal block: exactly what k_strong with the ghosts marked 'fixed' would feed it)."""
    blk = Abase_rows[:, rows].tocsr()
    blk.sort_indices()
    lv = D.amg.host_hierarchy(blk.indptr, blk.indices, blk.data, fixed[rows], coarse_max=max(8, rows.size // 3))
    if not lv[0]["nc"]:
        return -np.ones(rows.size, dtype=np.int64), 0
    return lv[0]["agg"], lv[0]["nc"]


def build_level(Abase, Apat, owner, fixed, nranks, net):
    """One distributed coarsening step of the BASE operator.  Apat carries the frozen structural pattern of the level (the
    library stores the base operator on it, explicit zeros included; scipy prunes them, so the ghost sets are taken from
    the current operator here).  Returns per-rank transfer data + the coarse ownership."""
    n = Abase.shape[0]
    Abase = Abase.tocsr()
    Apat = Apat.tocsr()
    ranks = []
    counts = []
    for r in range(nranks):
        rows = np.nonzero(owner == r)[0]
        Arows = Abase[rows]                                   # the rows this rank owns (global column ids)
        agg_loc, nc = local_aggregates(Arows, rows, fixed)
        ranks.append(dict(rows=rows, Arows=Arows, agg_loc=agg_loc, nc=nc))
        counts.append(nc)
    offs = np.concatenate([[0], np.cumsum(counts)])           # allgather of one int per rank
    agg = -np.ones(n, dtype=np.int64)                          # what each rank knows: own entries + ghosts after message 1
    for r, R in enumerate(ranks):
        R["agg_own"] = np.where(R["agg_loc"] >= 0, R["agg_loc"] + offs[r], -1)
        agg[R["rows"]] = R["agg_own"]
    nc_tot = int(offs[-1])
    owner_c = np.repeat(np.arange(nranks), counts)
    # message 1: aggregate ids of my ghost fine nodes, from their owners
    for r, R in enumerate(ranks):
        cols = np.unique(Apat[R["rows"]].indices)
        ghosts = cols[owner[cols] != r]
        R["ghosts"] = ghosts
        for q in np.unique(owner[ghosts]):
            net.send("1 ghost aggregate ids", q, r, np.sum(owner[ghosts] == q))
        R["agg_ghost"] = dict(zip(ghosts.tolist(), agg[ghosts].tolist()))
    # prolongator rows of the owned fine rows: P_i = (1-w) e_agg(i) - w/d_i sum_j a_ij e_agg(j)   (fp32 like the library)
    for r, R in enumerate(ranks):
        rows, Arows = R["rows"], R["Arows"].tocsr()
        lookup = agg.copy()                                    # legal: only own + ghost entries are read below
        data, indices, indptr = [], [], [0]
        d = Abase.diagonal()[rows]
        for li, i in enumerate(rows):
            acc = {}
            if not fixed[i]:
                if lookup[i] >= 0:
                    acc[lookup[i]] = 1.0 - OMEGA
                for k in range(Arows.indptr[li], Arows.indptr[li + 1]):
                    j, a = Arows.indices[k], Arows.data[k]
                    if j == i or a == 0.0 or lookup[j] < 0:
                        continue
                    acc[lookup[j]] = acc.get(lookup[j], 0.0) + (-OMEGA * a / d[li])
            for J in sorted(acc):
                indices.append(J)
                data.append(acc[J])
            indptr.append(len(indices))
        R["P"] = sp.csr_matrix((f32(data), indices, indptr), shape=(rows.size, nc_tot))
    # message 2: prolongator rows of my ghost fine nodes (needed by A*P of my rows and by my restriction rows)
    Pglob = sp.vstack([R["P"] for R in ranks]).tocsr()
    perm = np.concatenate([R["rows"] for R in ranks])
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    Pglob = Pglob[inv]                                         # global prolongator (rows in global numbering): reference only
    for r, R in enumerate(ranks):
        for q in np.unique(owner[R["ghosts"]]):
            g = R["ghosts"][owner[R["ghosts"]] == q]
            net.send("2 ghost prolongator rows", q, r, Pglob[g].nnz)
        R["P_ghost"] = {int(g): Pglob[int(g)] for g in R["ghosts"]}
    return ranks, owner_c, nc_tot, Pglob


def dist_galerkin(Aop, ranks, owner, owner_c, nc_tot, nranks, net, Pglob):
    """Coarse rows of every rank: C(I,:) = sum_i P(i,I) (A P)(i,:), i owned or ghost.  (A P)(i,:) of an owned row needs the
    prolongator rows of ghost columns (message 2); (A P)(i,:) of a ghost row comes from its owner (messages 3 + 4)."""
    Aop = Aop.tocsr()
    AP_rows = {}
    for r, R in enumerate(ranks):
        rows = R["rows"]
        cols = np.unique(Aop[rows].indices)
        # local prolongator slab: own rows + ghost rows, nothing else
        need = np.concatenate([rows, cols[owner[cols] != r]])
        Pslab = sp.lil_matrix((Aop.shape[0], nc_tot))
        Pslab[rows] = R["P"]
        for g in cols[owner[cols] != r]:
            Pslab[int(g)] = R["P_ghost"][int(g)]
        AP_rows[r] = (Aop[rows] @ Pslab.tocsr()).tocsr()       # rows of A*P for the owned fine rows
        assert set(need.tolist()) >= set(cols.tolist())
    coarse_rows = []
    for r, R in enumerate(ranks):
        own_c = np.nonzero(owner_c == r)[0]
        # fine rows with a prolongator entry in one of my aggregates: my rows + some ghost rows
        touch_own = R["P"][:, own_c]
        ghost_touch = [g for g in R["ghosts"] if R["P_ghost"][int(g)][:, own_c].nnz]
        for q in np.unique(owner[np.array(ghost_touch, dtype=np.int64)]) if ghost_touch else []:
            gq = [g for g in ghost_touch if owner[g] == q]
            rowpos = {int(i): k for k, i in enumerate(ranks[q]["rows"])}
            items = sum(AP_rows[q][rowpos[int(g)]].nnz for g in gq)
            net.send("3 A*P pattern of ghost rows (setup)", q, r, items)
            net.send("4 A*P values of ghost rows (per operator)", q, r, items)
        C = (touch_own.T @ AP_rows[r]).tocsr()
        for g in ghost_touch:
            q = owner[g]
            rowpos = int(np.searchsorted(ranks[q]["rows"], g))
            C = C + R["P_ghost"][int(g)][:, own_c].T @ AP_rows[q][rowpos]
        coarse_rows.append((own_c, C.tocsr()))
    Cglob = sp.lil_matrix((nc_tot, nc_tot))
    for own_c, C in coarse_rows:
        Cglob[own_c] = C
    Cglob = Cglob.tocsr()
    ref = (Pglob.T @ Aop @ Pglob).tocsr()
    err = abs(Cglob - ref).max() / abs(ref).max()
    return Cglob, err


class DistLevel:
    def __init__(self, A, owner, ranks, Pglob, owner_c, nranks, net):
        self.A, self.owner, self.ranks, self.P, self.owner_c, self.nranks, self.net = A.tocsr(), owner, ranks, Pglob.tocsr(), owner_c, nranks, net
        self.dinv = 1.0 / self.A.diagonal()
        self.rows = [R["rows"] for R in ranks]
        self.ghost = []
        for r in range(nranks):
            cols = np.unique(self.A[self.rows[r]].indices)
            self.ghost.append(cols[owner[cols] != r])

    def halo(self, vec_parts, kind, ghosts=None):
        """Every rank receives the values of its ghost nodes from their owners; returns per-rank {global id: value}."""
        out = []
        full = np.zeros(self.A.shape[0])
        for r in range(self.nranks):
            full[self.rows[r]] = vec_parts[r]
        for r in range(self.nranks):
            g = self.ghost[r] if ghosts is None else ghosts[r]
            for q in np.unique(self.owner[g]):
                self.net.send(kind, q, r, np.sum(self.owner[g] == q))
            out.append(full)                                   # (ranks only read own + ghost entries of it below)
        return out

    def spmv_rows(self, r, xfull):
        return self.A[self.rows[r]] @ xfull


def dist_vcycle(levels, tail_ops, tail_Ps, l, b_parts, net):
    L = levels[l]
    nr = L.nranks
    x = [OMEGA * L.dinv[L.rows[r]] * b_parts[r] for r in range(nr)]
    xf = L.halo(x, "5 halo(x) before the residual (per cycle)")
    t = [b_parts[r] - L.spmv_rows(r, xf[r]) for r in range(nr)]
    tf = L.halo(t, "6 halo(t) before the restriction (per cycle)")
    own_c = [np.nonzero(L.owner_c == r)[0] for r in range(nr)]
    bc = [np.asarray(L.P[:, own_c[r]].T @ tf[r]).ravel() for r in range(nr)]      # P(:,I) only touches own + ghost fine rows
    if l + 1 < len(levels):
        xc = dist_vcycle(levels, tail_ops, tail_Ps, l + 1, bc, net)
    else:                                                                          # replicated tail
        bfull = np.zeros(L.P.shape[1])
        for r in range(nr):
            bfull[own_c[r]] = bc[r]
            net.send("9 allgather of the tail's right-hand side (per cycle)", r, (r + 1) % nr, own_c[r].size * (nr - 1))
        xfull = amg_ref.vcycle(tail_ops, tail_Ps, bfull, OMEGA)
        xc = [xfull[own_c[r]] for r in range(nr)]
    # prolongation: coarse values of own aggregates + of the aggregates of ghost fine neighbours
    cfull = np.zeros(L.P.shape[1])
    for r in range(nr):
        cfull[own_c[r]] = xc[r]
    for r in range(nr):
        need = np.unique(L.P[L.rows[r]].indices)
        for q in np.unique(L.owner_c[need]):
            net.send("7 halo(xc) before the prolongation (per cycle)", q, r, np.sum(L.owner_c[need] == q))
    x = [x[r] + np.asarray(L.P[L.rows[r]] @ cfull).ravel() for r in range(nr)]
    xf = L.halo(x, "8 halo(x) before the smoothing sweep (per cycle)")
    return [x[r] + OMEGA * L.dinv[L.rows[r]] * (b_parts[r] - L.spmv_rows(r, xf[r])) for r in range(nr)]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    nx, ny, nranks = int(args[0]), int(args[1]), int(args[2])
    tail = int(sys.argv[sys.argv.index("--tail") + 1]) if "--tail" in sys.argv else 3000
    mesh, MV, A, b, isD = PR.operators(nx, ny)
    MV = MV.tocsr(); MV.sort_indices(); A = A.tocsr()
    n = A.shape[0]
    lib = D._lib.load()
    nodes = np.ascontiguousarray(mesh.nodes.T)
    owner = np.zeros(n, dtype=np.int32)
    lib.ddps_partition_nodes(nodes.ctypes.data_as(D._lib.c_double_p), n, nranks, owner.ctypes.data_as(D._lib.c_int32_p))
    net = Net()

    # ---- setup: distributed levels while they are large, then a replicated tail built by the library ----
    levels, Pglobs, ops = [], [], [A]
    base, own, fixed = MV, owner.astype(np.int64), isD
    while base.shape[0] > tail:
        ranks, owner_c, nc_tot, Pglob = build_level(base, ops[-1], own, fixed, nranks, net)
        Cop, err = dist_galerkin(ops[-1], ranks, own, owner_c, nc_tot, nranks, net, Pglob)
        print(f"level {len(levels)}: {base.shape[0]} -> {nc_tot} rows, distributed Galerkin vs P^T A P: {err:.1e}", flush=True)
        assert err < 1e-12
        levels.append(DistLevel(ops[-1], own, ranks, Pglob, owner_c, nranks, net))
        Pglobs.append(Pglob)
        ops.append(Cop)
        base = (Pglob.T @ base @ Pglob).tocsr()
        base.sort_indices()
        own, fixed = owner_c.astype(np.int64), np.zeros(nc_tot, dtype=bool)
    tail_lv = D.amg.host_hierarchy(base.indptr, base.indices, base.data, np.zeros(base.shape[0], dtype=bool))
    tail_ops, tail_Ps = amg_ref.operator_levels(ops[-1], tail_lv)
    for r in range(nranks):
        net.send("9 allgather of the tail's operator (per operator)", r, (r + 1) % nranks, ops[-1].nnz // nranks * (nranks - 1))
    print("hierarchy:", [L.A.shape[0] for L in levels], "| replicated tail:", [o.shape[0] for o in tail_ops], flush=True)

    # ---- (ii) one distributed cycle == the global cycle on the same hierarchy ----
    all_ops = ops[:-1] + tail_ops
    all_Ps = Pglobs + tail_Ps
    rng = np.random.default_rng(0)
    r0 = rng.standard_normal(n)
    parts = [r0[levels[0].rows[q]] for q in range(nranks)]
    zd = dist_vcycle(levels, tail_ops, tail_Ps, 0, parts, net)
    z = np.zeros(n)
    for q in range(nranks):
        z[levels[0].rows[q]] = zd[q]
    zg = amg_ref.vcycle(all_ops, all_Ps, r0, OMEGA)
    print("distributed cycle vs global cycle:", float(np.max(np.abs(z - zg)) / np.max(np.abs(zg))), flush=True)
    assert np.max(np.abs(z - zg)) <= 1e-12 * np.max(np.abs(zg))
    cycle_msgs = {k: tuple(v) for k, v in net.log.items()}

    # ---- (iii) CG iterations: distributed hierarchy vs the single-GPU hierarchy ----
    _, it_dist = amg_ref.pcg(A, b, lambda r: amg_ref.vcycle(all_ops, all_Ps, r, OMEGA))
    ref_lv = D.amg.host_hierarchy(MV.indptr, MV.indices, MV.data, isD)
    rops, rPs = amg_ref.operator_levels(A, ref_lv)
    _, it_ref = amg_ref.pcg(A, b, lambda r: amg_ref.vcycle(rops, rPs, r, OMEGA))
    print(f"CG iterations: rank-local aggregation on {nranks} ranks {it_dist}, single-GPU aggregation {it_ref}", flush=True)
    print("messages (count, payload items) for the setup, one operator and ONE cycle:")
    for k in sorted(cycle_msgs):
        print(f"  {k:58s} {cycle_msgs[k]}")


if __name__ == "__main__":
    main()
