#!/usr/bin/env python
"""CPU prototype for the NEXT step ("distributed hierarchy"): how many CG iterations does the multigrid
cycle lose when the finest-level aggregates may not cross the rank boundaries of the node partition (rank-local aggregation,
prolongator smoothing and cycle still global; the levels below are built as today and would be replicated)?
Test infrastructure, not product.   python tools/amg_mgpu_prototype.py 1024 512"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import amg_prototype as P  # noqa: E402
import amg_ref  # noqa: E402
import ddps_b200 as D  # noqa: E402

nx, ny = int(sys.argv[1]), int(sys.argv[2])
mesh, MV, A, b, isD = P.operators(nx, ny)
MV = MV.tocsr(); MV.sort_indices(); A = A.tocsr()
n = A.shape[0]
lib = D._lib.load()
nodes = np.ascontiguousarray(mesh.nodes.T)
omega = 2.0 / 3.0


def cycle_its(levels0_P, MVc):
    """levels0_P: finest-level prolongator; the rest of the hierarchy from the host library on the coarse base operator."""
    rest = D.amg.host_hierarchy(MVc.indptr, MVc.indices, MVc.data, np.zeros(MVc.shape[0], dtype=bool))
    Ps = [levels0_P] + [m[1] for m in amg_ref.hierarchy_matrices(rest)]
    ops = [A]
    for Pm in Ps[:-1]:
        ops.append((Pm.T @ ops[-1] @ Pm).tocsr())
    x, it = amg_ref.pcg(A, b, lambda r: amg_ref.vcycle(ops, Ps, r))
    return it, [o.shape[0] for o in ops]


def smoothed_P(agg, nc):
    i = np.nonzero(agg >= 0)[0]
    T = sp.csr_matrix((np.ones(i.size), (i, agg[i])), shape=(n, nc))
    S = sp.diags((~isD).astype(float)) @ (sp.identity(n) - omega * sp.diags(1.0 / MV.diagonal()) @ MV)
    Pm = (S @ T).tocsr()
    Pm.data = Pm.data.astype(np.float32).astype(np.float64)
    return Pm


ref = D.amg.host_hierarchy(MV.indptr, MV.indices, MV.data, isD)
P0 = amg_ref.hierarchy_matrices(ref)[0][1]
MVc = (P0.T @ MV @ P0).tocsr(); MVc.sort_indices()
print("global aggregation:", cycle_its(P0, MVc), flush=True)
for nparts in (2, 4, 8):
    owner = np.zeros(n, dtype=np.int32)
    lib.ddps_partition_nodes(nodes.ctypes.data_as(D._lib.c_double_p), n, nparts, owner.ctypes.data_as(D._lib.c_int32_p))
    agg = -np.ones(n, dtype=np.int64)
    nc = 0
    for p in range(nparts):
        idx = np.nonzero(owner == p)[0]
        MVp = MV[idx][:, idx].tocsr(); MVp.sort_indices()          # couplings to other ranks are invisible to the aggregation
        lv = D.amg.host_hierarchy(MVp.indptr, MVp.indices, MVp.data, isD[idx], coarse_max=10 ** 9)   # one level: aggregates only
        lv = D.amg.host_hierarchy(MVp.indptr, MVp.indices, MVp.data, isD[idx], coarse_max=max(8, int(0.5 * idx.size)))
        a = lv[0]["agg"]
        agg[idx] = np.where(a >= 0, a + nc, -1)
        nc += lv[0]["nc"]
    Pm = smoothed_P(agg, nc)
    MVc = (Pm.T @ MV @ Pm).tocsr(); MVc.sort_indices()
    print(f"rank-local aggregation, {nparts} parts:", cycle_its(Pm, MVc), flush=True)
