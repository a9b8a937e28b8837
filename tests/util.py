import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEYS = ("rec_mode", "Vbddata", "ubddata", "vbddata", "lam", "delta", "tau_p", "tau_n", "mu_p", "mu_n", "C", "tol")


def load_mesh_arrays(name):
    z = np.load(os.path.join(GOLDEN, f"mesh_{name}.npz"))
    return z["nodes"], z["edges"], z["elements"]


def gpu_mesh(name):
    import ddps_b200 as D
    return D.Mesh(*load_mesh_arrays(name))


def oracle_mesh(name_or_mesh):
    import ddps_oracle as O
    if isinstance(name_or_mesh, str):
        return O.Mesh(*load_mesh_arrays(name_or_mesh))
    return O.Mesh(name_or_mesh.nodes, name_or_mesh.edges, name_or_mesh.elements)


def golden(name):
    return np.load(os.path.join(GOLDEN, f"{name}.npz"))


def golden_csr(z, prefix, n):
    import scipy.sparse as sp
    return sp.csr_matrix((z[prefix + "_data"], z[prefix + "_indices"], z[prefix + "_indptr"]), shape=(n, n))


def rel_max(a, b):
    """CSR/vector parity metric of SURVEY Q6: max|a-b| / max|b|."""
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def csr_rel(A, B):
    """max|A-B|/max|B| over the union pattern (explicit zeros irrelevant)."""
    D = (A - B).tocoo()
    num = np.max(np.abs(D.data)) if D.nnz else 0.0
    return float(num / np.max(np.abs(B.data)))


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b))


def four_tag(problem):
    out = dict(problem)
    for k in ("Vbddata", "ubddata", "vbddata"):
        tags, kinds, f = problem[k]
        out[k] = [tags[:4], kinds[:4], f[:4]]
    return out


def smooth_fields(mesh, seed=0):
    """Seeded smooth positive test fields (V of either sign, u,v > 0)."""
    rng = np.random.default_rng(seed)
    x, y = mesh.nodes
    c = rng.uniform(-1, 1, size=8)
    V = 0.8 * np.sin(c[0] * 3 * x + c[1]) * np.cos(c[2] * 2 * y) + 0.3 * c[3] * x
    u = np.exp(0.5 * np.sin(c[4] * 2 * x + y))
    v = np.exp(-0.4 * np.cos(c[5] * x - 2 * y))
    return V, u, v


def fan_mesh(nfan=14, rings=3):
    """Unstructured test mesh with a high-degree centre node (nfan triangles around it, > 8 so the kernels'
    long-row / many-incidence paths run) surrounded by `rings` rings; boundary edges tagged 1 (lower half) / 2."""
    import ddps_b200 as D
    pts = [(0.0, 0.0)]
    for r in range(1, rings + 1):
        for k in range(nfan):
            a = 2 * np.pi * (k + 0.5 * (r % 2)) / nfan
            pts.append((r * np.cos(a) / rings, r * np.sin(a) / rings))
    nodes = np.array(pts).T
    idx = lambda r, k: 1 + (r - 1) * nfan + (k % nfan)          # 0-based id of ring r (1..), position k
    tris = []
    for k in range(nfan):
        tris.append((0, idx(1, k), idx(1, k + 1)))
    for r in range(1, rings):
        for k in range(nfan):
            a, b = idx(r, k), idx(r, k + 1)
            if r % 2 == 1:      # outer ring shifted by half a step relative to the inner one
                c, d = idx(r + 1, k), idx(r + 1, k + 1)
                tris.append((a, c, b))
                tris.append((b, c, d))
            else:
                c, d = idx(r + 1, k - 1), idx(r + 1, k)
                tris.append((a, c, d))
                tris.append((a, d, b))
    el = np.zeros((5, len(tris)), dtype=np.int64)
    for e, t in enumerate(tris):
        q = nodes[:, list(t)]
        area = (q[0, 1] - q[0, 2]) * (q[1, 2] - q[1, 0]) - (q[1, 1] - q[1, 2]) * (q[0, 2] - q[0, 0])
        t = t if area > 0 else (t[0], t[2], t[1])
        el[0:3, e] = np.array(t) + 1
    el[3:] = 2
    edges = []
    for k in range(nfan):
        a, b = idx(rings, k) + 1, idx(rings, k + 1) + 1
        tag = 1 if k < nfan // 2 else 2
        edges.append((a, b, 1, tag))
    return D.Mesh(nodes, np.array(edges, dtype=np.int64).T.copy(), el)
