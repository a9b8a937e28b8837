"""numpy/scipy restatement of the multigrid preconditioner (test infrastructure): the V(1,1) cycle with damped-Jacobi
smoothing, Galerkin operators of the current operator on the frozen prolongators, exact coarsest solve -- the
device kernels of csrc/amg.cu are checked against this."""
import numpy as np
import scipy.sparse as sp


def hierarchy_matrices(levels):
    """scipy views of ddps_b200.amg.host_hierarchy() / Session.amg_hierarchy(): list of (A_base or None, P or None)."""
    out = []
    for lev in levels:
        n = lev["n"]
        A = sp.csr_matrix((lev["val"], lev["col"], lev["rowptr"]), shape=(n, n)) if "val" in lev else None
        P = sp.csr_matrix((lev["Pval"], lev["Pcol"], lev["Pptr"]), shape=(n, lev["nc"])) if lev["nc"] else None
        out.append((A, P))
    return out


def operator_levels(A, levels):
    """Galerkin hierarchy of the operator A on the frozen prolongators."""
    mats = hierarchy_matrices(levels)
    ops = [A.tocsr()]
    for (_, P) in mats[:-1]:
        ops.append((P.T @ ops[-1] @ P).tocsr())
    return ops, [P for (_, P) in mats]


def vcycle(ops, Ps, r, omega=2.0 / 3.0, level=0, fp32_min=None):
    """fp32_min: levels with more rows than this keep an fp32-rounded copy of their matrix for the two SpMVs of the
    cycle (what the device does for the large levels; the inverse diagonal stays FP64)."""
    A = ops[level]
    if Ps[level] is None:
        return np.linalg.solve(A.toarray(), r)
    dinv = 1.0 / A.diagonal()
    As = A
    if fp32_min is not None and A.shape[0] > fp32_min:
        As = A.copy()
        As.data = As.data.astype(np.float32).astype(np.float64)
    x = omega * dinv * r
    t = r - As @ x
    xc = vcycle(ops, Ps, Ps[level].T @ t, omega, level + 1, fp32_min)
    x = x + Ps[level] @ xc
    return x + omega * dinv * (r - As @ x)


def pcg(A, b, M, rtol=1e-13, maxit=2000):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = r @ z
    bb = np.sqrt(b @ b)
    for it in range(1, maxit + 1):
        q = A @ p
        al = rz / (p @ q)
        x += al * p
        r -= al * q
        if np.sqrt(r @ r) <= rtol * bb:
            return x, it
        z = M(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxit


def cycle_omega(base):
    """Damping of the cycle's Jacobi sweeps as the library chooses it (csrc/amg.cu amg_cycle_omega): min(0.8, 1.6 / rho_G) with the
    Gershgorin bound rho_G = max_i sum_j |a_ij| / a_ii of the BASE operator (2 for the stiffness matrix of an acute mesh)."""
    import numpy as np
    A = base.tocsr()
    d = A.diagonal()
    rs = np.asarray(abs(A).sum(axis=1)).ravel()
    ok = d > 0
    return min(0.8, 1.6 / float(np.max(rs[ok] / d[ok])))

