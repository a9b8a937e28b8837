"""CPU test of the N>1 host logic: two gloo ranks each derive their local mesh + halo plan through
the host-only C-ABI entry point (ddps_partition_plan = the code ddps_mesh_set runs), exchange their
owned boundary values exactly as the library's halo exchange does (send lists / ghost blocks), and
verify that every ghost slot receives the right global value.  Also emulates a partitioned SpMV."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def plan(D, mesh, nparts, rank):
    from ddps_b200 import _lib
    lib = _lib.load()
    nodes = np.ascontiguousarray(mesh.nodes.T)
    el = np.ascontiguousarray(mesh.elements.T)
    counts = np.zeros(4, dtype=np.int64)
    l2g = np.zeros(mesh.nq, dtype=np.int64)
    peers = np.zeros(nparts, dtype=np.int32)
    sc = np.zeros(nparts, dtype=np.int64)
    rc_ = np.zeros(nparts, dtype=np.int64)
    sg = np.zeros(mesh.nq, dtype=np.int64)
    rc = lib.ddps_partition_plan(_lib.dptr(nodes), mesh.nq, _lib.iptr(el), mesh.nme, nparts, rank, _lib.iptr(counts),
                                 _lib.iptr(l2g), peers.ctypes.data_as(_lib.c_int32_p), _lib.iptr(sc), _lib.iptr(rc_), _lib.iptr(sg))
    assert rc == 0
    n_own, n_ghost, nme_loc, nn = (int(c) for c in counts)
    return dict(n_own=n_own, n_ghost=n_ghost, nme=nme_loc, l2g=l2g[:n_own + n_ghost], peers=peers[:nn].tolist(),
                send_counts=sc[:nn].tolist(), recv_counts=rc_[:nn].tolist(), send_glob=sg[:int(sc[:nn].sum())])


def _worker(rank, world, port, which):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ddps_b200 as D
    from util import gpu_mesh
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        mesh = gpu_mesh("F2") if which == "F2" else D.structured_mesh(24, 12, 0, 2, 0, 1)
        P = plan(D, mesh, world, rank)
        x_glob = np.sin(np.arange(mesh.nq) * 0.37) + 2.0       # a global nodal field every rank can evaluate
        local = np.full(P["n_own"] + P["n_ghost"], np.nan)
        local[:P["n_own"]] = x_glob[P["l2g"][:P["n_own"]]]
        # halo exchange: pack send list, send to each peer, receive into the contiguous ghost block
        reqs, so, ro = [], 0, 0
        recv_bufs = []
        for j, peer in enumerate(P["peers"]):
            sbuf = torch.from_numpy(x_glob[P["send_glob"][so:so + P["send_counts"][j]]].copy())
            rbuf = torch.empty(P["recv_counts"][j], dtype=torch.float64)
            reqs.append(dist.isend(sbuf, peer))
            reqs.append(dist.irecv(rbuf, peer))
            recv_bufs.append((ro, rbuf))
            so += P["send_counts"][j]
            ro += P["recv_counts"][j]
        for r in reqs:
            r.wait()
        for off, rbuf in recv_bufs:
            local[P["n_own"] + off:P["n_own"] + off + rbuf.numel()] = rbuf.numpy()
        assert not np.any(np.isnan(local))
        assert np.array_equal(local, x_glob[P["l2g"]])           # every ghost got its owner's value
        # ownership is a partition of the nodes
        own = torch.zeros(mesh.nq, dtype=torch.int32)
        own[torch.from_numpy(P["l2g"][:P["n_own"]])] = 1
        dist.all_reduce(own)
        assert int(own.min()) == 1 and int(own.max()) == 1
        # partitioned graph-Laplacian SpMV over the LOCAL triangles reproduces the global one on owned rows
        g2l = {int(g): i for i, g in enumerate(P["l2g"])}
        tri = mesh.elements[:3].T - 1
        y_glob = np.zeros(mesh.nq)
        for a, b, c in tri:
            for p, q in ((a, b), (b, c), (c, a)):
                y_glob[p] += x_glob[p] - x_glob[q]
                y_glob[q] += x_glob[q] - x_glob[p]
        y_loc = np.zeros(P["n_own"])
        ntri = 0
        for a, b, c in tri:
            if not any(g2l.get(int(t), 10 ** 9) < P["n_own"] for t in (a, b, c)):
                continue
            ntri += 1
            for p, q in ((a, b), (b, c), (c, a)):
                lp, lq = g2l[int(p)], g2l[int(q)]
                if lp < P["n_own"]:
                    y_loc[lp] += local[lp] - local[lq]
                if lq < P["n_own"]:
                    y_loc[lq] += local[lq] - local[lp]
        assert ntri == P["nme"]
        assert np.allclose(y_loc, y_glob[P["l2g"][:P["n_own"]]], rtol=0, atol=1e-12)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("which", ["structured", "F2"])
def test_halo_plan_world2_gloo(which):
    mp.spawn(_worker, args=(2, _free_port(), which), nprocs=2, join=True)


def test_plan_consistency_4_parts():
    """send list of p -> q equals the ghost block q holds for p, for every pair, without communication."""
    import ddps_b200 as D
    mesh = D.structured_mesh(20, 20)
    plans = [plan(D, mesh, 4, r) for r in range(4)]
    for p, P in enumerate(plans):
        so = 0
        for j, q in enumerate(P["peers"]):
            sent = P["send_glob"][so:so + P["send_counts"][j]]
            so += P["send_counts"][j]
            Q = plans[q]
            jj = Q["peers"].index(p)
            ro = sum(Q["recv_counts"][:jj])
            ghost = Q["l2g"][Q["n_own"] + ro:Q["n_own"] + ro + Q["recv_counts"][jj]]
            assert np.array_equal(sent, ghost)
    assert sum(P["n_own"] for P in plans) == mesh.nq
