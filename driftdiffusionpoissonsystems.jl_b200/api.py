"""Host-side mirror of the reference's public API for the hot path, on top of the C ABI.

Same names, argument order and error behaviour as the reference
(``/root/reference/src/DriftDiffusionPoissonSystems.jl``, cited ``src:N``):

    solve_ddpe(mesh, rec_mode, Vbddata, ubddata, vbddata, lambda, delta, tau_p, tau_n, mu_p, mu_n, C, tol=1e-9) -> (V,u,v)      src:630
    solve_semlin_poisson(mesh, A, bddata, f, g, tol=1e-14) -> V                                                                src:895
    aquire_boundary(mesh, bddata) -> (neu_nodes1, neu_nodes2, gD, n)                                                          src:162

What runs where: this module evaluates the user's *spatial* closures ONCE on the host (lambda, mu at
centroids; C / h at edge midpoints; boundary closures at boundary nodes / edge midpoints) and uploads
the samples; everything per-iteration (assembly, Dirichlet elimination, PCG, Newton and Gummel
loops) runs on the GPU inside libddps_b200.so.  u-dependent right-hand sides must be registered
functors (functors.py) -- arbitrary closures cannot run on the device and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .functors import FUNCTOR_POLY, FUNCTOR_SINH, PolyRHS, SinhRHS, SpatialFunctor
from .mesh import Mesh

FIELD_V, FIELD_U, FIELD_W, FIELD_SCALAR = 0, 1, 2, 3
COEFF_LAMBDA2, COEFF_MU_P, COEFF_MU_N, COEFF_C_MID, COEFF_A11, COEFF_A22, COEFF_H_MID, COEFF_NEUMANN = range(8)
COEFF_POLY_MID0, COEFF_POLY_NODE0 = 16, 32
REC_MODES = {"zero": 0, "shockley": 1, ":zero": 0, ":shockley": 1}
MAT_BASE, MAT_OP = 0, 1
VEC_RHS, VEC_RESID, VEC_V, VEC_U, VEC_W, VEC_LIFT = range(6)


class DDPSError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ddps error {code}: {msg}")
        self.code = code


# --------------------------------------------------------------------------------------------
# host-side boundary preprocessing (src:162-218 and its inline copies src:532-544, 965-1013)
# --------------------------------------------------------------------------------------------
_SAMPLE_THREADS = max(1, min(16, (__import__("os").cpu_count() or 1)))
_SAMPLE_CHUNK = 1 << 20


def _sample2(fn, x, y):
    """Evaluate a scalar closure (x,y)->value on arrays (Julia's ``map``/dot call).  Large arrays are cut into chunks
    that a thread pool evaluates concurrently (vectorised closures spend their time in numpy ufuncs, which release the
    GIL): the one-time host sampling of arbitrary closures uses all host cores."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if x.size >= 4 * _SAMPLE_CHUNK and _SAMPLE_THREADS > 1:
        from concurrent.futures import ThreadPoolExecutor
        xf, yf = x.reshape(-1), y.reshape(-1)
        out = np.empty(xf.size)
        cuts = list(range(0, xf.size, _SAMPLE_CHUNK))

        def work(o):
            out[o:o + _SAMPLE_CHUNK] = _sample2_serial(fn, xf[o:o + _SAMPLE_CHUNK], yf[o:o + _SAMPLE_CHUNK])

        with ThreadPoolExecutor(_SAMPLE_THREADS) as ex:
            list(ex.map(work, cuts))
        return out.reshape(x.shape)
    return _sample2_serial(fn, x, y)


def _sample2_serial(fn, x, y):
    try:
        out = np.asarray(fn(x, y), dtype=np.float64)
        if out.shape == x.shape:
            return np.ascontiguousarray(out)
        if out.shape == ():
            return np.full(x.shape, float(out))
    except Exception:
        pass
    return np.array([float(fn(a, b)) for a, b in zip(x.ravel(), y.ravel())]).reshape(x.shape)


def _split_bddata(bddata):
    tags, kinds, fcts = bddata
    if not (len(tags) == len(kinds) == len(fcts)):
        raise ValueError("bddata must be a 3 x n array: tags / 'D' or 'N' / closures")
    neu = [j for j, k in enumerate(kinds) if k == "N"]
    diri = [j for j, k in enumerate(kinds) if k == "D"]
    if len(neu) + len(diri) != len(tags):      # src:169-171, 971-973
        raise ValueError("Could not understand boundary file. Not all parts of the boundary have BC assigned.")
    return tags, fcts, neu, diri


def dirichlet_nodes(mesh: Mesh, bddata):
    """Dirichlet node ids ``n`` (1-based, order of first appearance along the listed segments) and
    their values ``gD``.  Follows src:199-211: per 'D' column the nodes of the segment's edges in
    (n1,n2)-interleaved order without repeats, values from that column's closure, then duplicate
    (node,value) rows collapsed."""
    tags, fcts, _, diri = _split_bddata(bddata)
    nodes_acc, vals_acc = [], []
    for j in diri:
        sel = mesh.edges[3] == tags[j]
        seq = mesh.edges[0:2, sel].T.reshape(-1)
        _, first = np.unique(seq, return_index=True)
        seg = seq[np.sort(first)]
        if seg.size != int(sel.sum()) + 1:
            raise ValueError(f"boundary segment {tags[j]} is not a single open polyline "
                             "(the reference's value list would not line up, src:204-206)")
        nodes_acc.append(seg)
        vals_acc.append(_sample2(fcts[j], mesh.nodes[0, seg - 1], mesh.nodes[1, seg - 1]))
    if not nodes_acc:
        return np.zeros(0, dtype=np.int64), np.zeros(0)
    nd = np.concatenate(nodes_acc)
    vl = np.concatenate(vals_acc)
    keep, seen = [], set()
    for k, key in enumerate(zip(nd.tolist(), vl.tolist())):
        if key not in seen:
            seen.add(key)
            keep.append(k)
    return nd[keep].astype(np.int64), vl[keep].astype(np.float64)


def neumann_edges(mesh: Mesh, bddata):
    """End nodes of the Neumann edges and the edge loads len(E)/2*gN(midpoint) (src:976-991)."""
    tags, fcts, neu, _ = _split_bddata(bddata)
    a_all, b_all, load_all = [], [], []
    for j in neu:
        sel = mesh.edges[3] == tags[j]
        a, b = mesh.edges[0, sel], mesh.edges[1, sel]
        xa, ya, xb, yb = mesh.nodes[0, a - 1], mesh.nodes[1, a - 1], mesh.nodes[0, b - 1], mesh.nodes[1, b - 1]
        gn = _sample2(fcts[j], (xa + xb) / 2, (ya + yb) / 2)
        a_all.append(a)
        b_all.append(b)
        load_all.append(np.sqrt((xa - xb) ** 2 + (ya - yb) ** 2) * gn / 2)
    if not a_all:
        z = np.zeros(0, dtype=np.int64)
        return z, z, np.zeros(0)
    return np.concatenate(a_all), np.concatenate(b_all), np.concatenate(load_all)


def aquire_boundary(mesh: Mesh, bddata):
    """src:162-218 -> (neu_nodes1, neu_nodes2, gD, n)."""
    n1, n2, _ = neumann_edges(mesh, bddata)
    n, gD = dirichlet_nodes(mesh, bddata)
    return n1, n2, gD, n


def _centroids(mesh: Mesh):
    a = mesh.elements[0:3] - 1
    X, Y = mesh.nodes
    return (X[a[0]] + X[a[1]] + X[a[2]]) / 3, (Y[a[0]] + Y[a[1]] + Y[a[2]]) / 3        # src:250-251


def _midpoints(mesh: Mesh):
    """[nme,3] arrays of the edge midpoints m12, m23, m31 (src:303-308)."""
    a = mesh.elements[0:3] - 1
    X, Y = mesh.nodes
    mx = np.stack([(X[a[0]] + X[a[1]]) / 2, (X[a[1]] + X[a[2]]) / 2, (X[a[2]] + X[a[0]]) / 2], axis=1)
    my = np.stack([(Y[a[0]] + Y[a[1]]) / 2, (Y[a[1]] + Y[a[2]]) / 2, (Y[a[2]] + Y[a[0]]) / 2], axis=1)
    return mx, my


def sample_ddpe(mesh: Mesh, Vbddata, ubddata, vbddata, lam, mu_p, mu_n, Cfun, device_functors=False):
    """One-time host setup of solve_ddpe: evaluate the spatial closures where the reference
    evaluates them (boundary nodes, centroids, edge midpoints) and return plain arrays.
    ``device_functors``: coefficient closures that are registered spatial functors (functors.SpatialFunctor) are NOT
    sampled here; the entry holds the functor itself and ``Session.upload_ddpe`` lets the library sample it on the device."""
    _, _, gD, n = aquire_boundary(mesh, Vbddata)                             # src:640
    nu, gDu = dirichlet_nodes(mesh, ubddata)                                 # src:532-544
    nv, gDv = dirichlet_nodes(mesh, vbddata)
    if len(gDu) != len(n) or len(gDv) != len(n):
        raise ValueError("DimensionMismatch: the u/v bddata must mark the same segments 'D' as the V bddata (src:563)")
    fns = dict(lambda2=lam, mu_p=mu_p, mu_n=mu_n, C_mid=Cfun)
    host = {k: f for k, f in fns.items() if not (device_functors and isinstance(f, SpatialFunctor))}
    out = dict(n=n, gD_V=gD, gD_u=gDu, gD_v=gDv)
    out.update({k: f for k, f in fns.items() if k not in host})
    if any(k in host for k in ("lambda2", "mu_p", "mu_n")):
        cx, cy = _centroids(mesh)
        if "lambda2" in host:
            out["lambda2"] = _sample2(lam, cx, cy) ** 2                      # src:254
        for k in ("mu_p", "mu_n"):
            if k in host:
                out[k] = _sample2(host[k], cx, cy)                           # src:467
    if "C_mid" in host:
        mx, my = _midpoints(mesh)
        out["C_mid"] = np.ascontiguousarray(_sample2(Cfun, mx, my))          # src:341-343
    return out


# --------------------------------------------------------------------------------------------
# Session: one library context (= one GPU)
# --------------------------------------------------------------------------------------------
class Session:
    """Owns a ``ddps_ctx``.  ``Session(device, comm=(id_bytes, rank, nranks))`` for multi-GPU."""

    def __init__(self, device=0, comm=None, **opts):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.ddps_ctx_create(C.byref(h), int(device))
        if rc != 0:
            raise DDPSError(rc, self.lib.ddps_last_error(None).decode())
        self.h = h
        self.mesh = None
        if comm is not None:
            ident, rank, nranks = comm
            self._check(self.lib.ddps_comm_init(self.h, ident, int(rank), int(nranks)))
        if opts:
            self.set_opts(**opts)

    # -- plumbing --
    def _check(self, rc):
        if rc != 0:
            raise DDPSError(rc, self.lib.ddps_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.ddps_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_opts(self, **kw):
        o = _lib.SolverOpts()
        self.lib.ddps_default_opts(C.byref(o))
        if hasattr(self, "_opts"):
            o = self._opts
        for k, v in kw.items():
            if not hasattr(o, k):
                raise TypeError(f"unknown solver option {k}")
            setattr(o, k, v)
        self._opts = o
        self._check(self.lib.ddps_set_opts(self.h, C.byref(o)))

    # -- uploads --
    def set_mesh(self, mesh: Mesh):
        nodes = np.ascontiguousarray(mesh.nodes.T, dtype=np.float64)         # [nq,2] == Julia 2 x nq column-major
        elements = np.ascontiguousarray(mesh.elements.T, dtype=np.int64)     # [nme,5] == Julia 5 x nme column-major
        self._check(self.lib.ddps_mesh_set(self.h, _lib.dptr(nodes), nodes.shape[0], _lib.iptr(elements),
                                           elements.shape[0]))
        self.mesh = mesh

    def set_mesh_local(self, part, with_elem_glob=True):
        """Per-rank mesh ingest (``ddps_mesh_set_local``): ``part`` as returned by ``mesh.structured_mesh_part`` -- nobody
        holds the global mesh.  ``self.mesh`` becomes the boundary-only Mesh the host preprocessing needs."""
        nodes = np.ascontiguousarray(part["nodes"], dtype=np.float64)
        el = np.ascontiguousarray(part["elements"], dtype=np.int64)
        ng = part["n_loc"] - part["n_own"]
        go = part["ghost_owner"]
        eg = part.get("elem_glob") if with_elem_glob else None
        self._check(self.lib.ddps_mesh_set_local(
            self.h, _lib.dptr(nodes), int(part["n_own"]), int(part["n_loc"]), _lib.iptr(part["node_glob"]),
            go.ctypes.data_as(_lib.c_int32_p) if ng else None, _lib.iptr(el), el.shape[0],
            _lib.iptr(eg) if eg is not None else None, int(part["nq_global"]), int(part["nme_global"])))
        self.mesh = part["boundary"]
        self._part = part

    def ddpe_fetch_owned(self):
        """V, u, v of the nodes this rank owns (order of ``part['own_glob']``)."""
        n = self._part["n_own"]
        V, u, v = np.empty(n), np.empty(n), np.empty(n)
        self._check(self.lib.ddps_ddpe_fetch_owned(self.h, _lib.dptr(V), _lib.dptr(u), _lib.dptr(v), n))
        return V, u, v

    def set_dirichlet(self, field, n, gD):
        n = np.ascontiguousarray(n, dtype=np.int64)
        gD = np.ascontiguousarray(gD, dtype=np.float64)
        self._check(self.lib.ddps_dirichlet_set(self.h, field, _lib.iptr(n), _lib.dptr(gD), n.size))

    def set_coeff(self, slot, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        self._check(self.lib.ddps_coeff_set(self.h, slot, _lib.dptr(arr), arr.size))

    def set_coeff_functor(self, slot, fn: SpatialFunctor):
        """Let the library sample a registered spatial functor on the device (no host sampling, no upload)."""
        p = np.ascontiguousarray(fn.params(), dtype=np.float64)
        self._check(self.lib.ddps_coeff_functor(self.h, slot, int(fn.kind), _lib.dptr(p), p.size))

    def set_coeff_any(self, slot, arr_or_functor):
        if isinstance(arr_or_functor, SpatialFunctor):
            self.set_coeff_functor(slot, arr_or_functor)
        else:
            self.set_coeff(slot, arr_or_functor)

    def set_functor(self, kind, params):
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._check(self.lib.ddps_functor_set(self.h, kind, _lib.dptr(p), p.size))

    # -- problem setup (host sampling of the spatial closures) --
    def setup_ddpe(self, Vbddata, ubddata, vbddata, lam, mu_p, mu_n, Cfun, device_functors=True):
        self.upload_ddpe(sample_ddpe(self.mesh, Vbddata, ubddata, vbddata, lam, mu_p, mu_n, Cfun, device_functors))

    def upload_ddpe(self, smp):
        """Hand the sampled arrays of ``sample_ddpe`` to the library (host buffers -> HBM)."""
        # the reference pairs gD_u positionally with V's node list n (src:563-564)
        self.set_dirichlet(FIELD_V, smp["n"], smp["gD_V"])
        self.set_dirichlet(FIELD_U, smp["n"], smp["gD_u"])
        self.set_dirichlet(FIELD_W, smp["n"], smp["gD_v"])
        self.set_coeff_any(COEFF_LAMBDA2, smp["lambda2"])   # arrays sampled on the host, or registered spatial functors
        self.set_coeff_any(COEFF_MU_P, smp["mu_p"])         # (sampled by the library on the device)
        self.set_coeff_any(COEFF_MU_N, smp["mu_n"])
        self.set_coeff_any(COEFF_C_MID, smp["C_mid"])
        self._ddpe_n = smp["n"]

    def ddpe_begin(self, rec_mode, delta, tau_p, tau_n, tol):
        self._check(self.lib.ddps_ddpe_begin(self.h, REC_MODES[str(rec_mode)], float(delta), float(tau_p),
                                             float(tau_n), float(tol)))

    def ddpe_step(self, nsteps=1, stop_at_tol=False):
        ctrl = np.zeros(max(nsteps, 1))
        done = C.c_int(0)
        self._check(self.lib.ddps_ddpe_step(self.h, int(nsteps), int(bool(stop_at_tol)), _lib.dptr(ctrl), C.byref(done)))
        return ctrl[:done.value]

    def ddpe_fetch(self, out=None):
        """Copy the current V, u, v to the host.  ``out`` = three caller-owned float64 arrays of length nq (e.g. pinned
        memory, like the caller-allocated outputs of the C ABI); default: fresh numpy arrays."""
        nq = self.mesh.nq
        V, u, v = out if out is not None else (np.empty(nq), np.empty(nq), np.empty(nq))
        for a in (V, u, v):
            if a.shape != (nq,) or a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("ddpe_fetch outputs must be contiguous float64 arrays of length nq")
        self._check(self.lib.ddps_ddpe_fetch(self.h, _lib.dptr(V), _lib.dptr(u), _lib.dptr(v)))
        return V, u, v

    def solve_ddpe(self, rec_mode, delta, tau_p, tau_n, tol):
        nq = self.mesh.nq
        V, u, v = np.empty(nq), np.empty(nq), np.empty(nq)
        st = _lib.Stats()
        rc = self.lib.ddps_solve_ddpe(self.h, REC_MODES[str(rec_mode)], float(delta), float(tau_p), float(tau_n),
                                      float(tol), _lib.dptr(V), _lib.dptr(u), _lib.dptr(v), C.byref(st))
        self._check(rc)
        return V, u, v, st.as_dict()

    def setup_semlin(self, A, bddata, f, g):
        mesh = self.mesh
        if not isinstance(f, (SinhRHS, PolyRHS)):
            raise TypeError("f must be a registered functor (SinhRHS / PolyRHS): arbitrary closures cannot run on "
                            "the device and this path has no CPU fallback")
        # g is evaluated on the device from the functor's parameters: make sure the closure the caller passed is
        # that derivative (the reference would silently use whatever g it is given, src:1046)
        xs, ys, us = np.array([0.1, -0.3, 0.7]), np.array([0.2, 0.5, -0.4]), np.array([0.3, -0.6, 1.1])
        if not np.allclose(np.asarray(g(xs, ys, us), dtype=float) + 0 * xs, f.derivative()(xs, ys, us), rtol=1e-10, atol=1e-12):
            raise ValueError("g must be the u-derivative of the registered functor f (use f.derivative())")
        gtag = mesh.elements[4].astype(np.float64)
        utag = np.unique(gtag)
        a11 = {t: float(A[0](t)) for t in utag}                              # src:901 (A is called on the geometric tag)
        a22 = {t: float(A[3](t)) for t in utag}                              # src:902
        self.set_coeff(COEFF_A11, np.array([a11[t] for t in gtag]))
        self.set_coeff(COEFF_A22, np.array([a22[t] for t in gtag]))
        n, gD = dirichlet_nodes(mesh, bddata)                                # src:1001-1013
        self.set_dirichlet(FIELD_SCALAR, n, gD)
        n1, n2, load = neumann_edges(mesh, bddata)                           # src:976-996
        nl = np.zeros(mesh.nq)
        np.add.at(nl, n1 - 1, load)
        np.add.at(nl, n2 - 1, load)
        self.set_coeff(COEFF_NEUMANN, nl)
        mx, my = _midpoints(mesh)
        if isinstance(f, SinhRHS):
            self.set_coeff(COEFF_H_MID, _sample2(f.h, mx, my))
            self.set_functor(FUNCTOR_SINH, [f.alpha, f.beta])
        else:
            for k, c in enumerate(f.coeffs):
                self.set_coeff(COEFF_POLY_MID0 + k, _sample2(c, mx, my))
                self.set_coeff(COEFF_POLY_NODE0 + k, _sample2(c, mesh.nodes[0], mesh.nodes[1]))
            self.set_functor(FUNCTOR_POLY, [len(f.coeffs) - 1])

    def solve_semlin(self, tol):
        V = np.empty(self.mesh.nq)
        st = _lib.Stats()
        rc = self.lib.ddps_solve_semlin(self.h, float(tol), _lib.dptr(V), C.byref(st))
        self._check(rc)
        return V, st.as_dict()

    # -- introspection --
    def stats(self):
        st = _lib.Stats()
        self._check(self.lib.ddps_get_stats(self.h, C.byref(st)))
        return st.as_dict()

    def history(self, kind=0):
        n = C.c_int64(0)
        self._check(self.lib.ddps_get_history(self.h, kind, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1))
        self._check(self.lib.ddps_get_history(self.h, kind, _lib.dptr(out), n.value, C.byref(n)))
        return out[:n.value]

    def get_csr(self, which=MAT_OP):
        """scipy CSR of the fixed pattern with the chosen values (0-based, explicit zeros kept)."""
        import scipy.sparse as sp
        nr, nnz = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.ddps_get_csr_size(self.h, C.byref(nr), C.byref(nnz)))
        rp = np.zeros(nr.value + 1, dtype=np.int64)
        col = np.zeros(nnz.value, dtype=np.int64)
        val = np.zeros(nnz.value)
        self._check(self.lib.ddps_get_csr(self.h, which, _lib.iptr(rp), _lib.iptr(col), _lib.dptr(val)))
        return sp.csr_matrix((val, col, rp), shape=(nr.value, self.mesh.nq))

    def get_vector(self, which):
        nr = C.c_int64(0)
        self._check(self.lib.ddps_get_csr_size(self.h, C.byref(nr), None))
        out = np.zeros(nr.value)
        self._check(self.lib.ddps_get_vector(self.h, which, _lib.dptr(out), out.size))
        return out

    def get_coeff(self, slot):
        n = C.c_int64(0)
        self._check(self.lib.ddps_get_coeff(self.h, int(slot), None, 0, C.byref(n)))
        out = np.zeros(n.value)
        self._check(self.lib.ddps_get_coeff(self.h, int(slot), _lib.dptr(out), out.size, C.byref(n)))
        return out

    def debug_assemble_base(self, semlin=False):
        self._check(self.lib.ddps_debug_assemble_base(self.h, int(semlin)))

    def debug_assemble_newton(self, semlin, delta, V, u=None, v=None):
        res = C.c_double(0)
        V = np.ascontiguousarray(V, dtype=np.float64)
        pu = _lib.dptr(np.ascontiguousarray(u, dtype=np.float64)) if u is not None else None
        pv = _lib.dptr(np.ascontiguousarray(v, dtype=np.float64)) if v is not None else None
        self._check(self.lib.ddps_debug_assemble_newton(self.h, int(semlin), float(delta), _lib.dptr(V), pu, pv, C.byref(res)))
        return res.value

    def debug_assemble_uv(self, which, rec_mode, tau_p, tau_n, V, u0, v0):
        V, u0, v0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (V, u0, v0))
        self._check(self.lib.ddps_debug_assemble_uv(self.h, which, REC_MODES[str(rec_mode)], float(tau_p), float(tau_n),
                                                    _lib.dptr(V), _lib.dptr(u0), _lib.dptr(v0)))

    def debug_spmv(self, which, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.get_vector_len())
        self._check(self.lib.ddps_debug_spmv(self.h, which, _lib.dptr(x), _lib.dptr(y)))
        return y

    def get_vector_len(self):
        nr = C.c_int64(0)
        self._check(self.lib.ddps_get_csr_size(self.h, C.byref(nr), None))
        return nr.value

    def debug_pcg(self, rhs, x0=None, rtol=1e-12, max_it=100000):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.zeros_like(rhs) if x0 is None else np.array(x0, dtype=np.float64)
        it, rr = C.c_int64(0), C.c_double(0)
        self._check(self.lib.ddps_debug_pcg(self.h, _lib.dptr(rhs), _lib.dptr(x), float(rtol), int(max_it), C.byref(it), C.byref(rr)))
        return x, it.value, rr.value

    # -- multigrid preconditioner (opts.precond = 1) --
    def amg_levels(self):
        nl = C.c_int64(0)
        sizes = np.zeros(16, dtype=np.int64)
        self._check(self.lib.ddps_amg_levels(self.h, C.byref(nl), _lib.iptr(sizes)))
        return [int(v) for v in sizes[:nl.value]]

    def amg_hierarchy(self):
        """The hierarchy as it lives on the device: per level a dict with n, nnz, nc, rowptr, col (CSR pattern, sorted
        columns) and, except on the coarsest level, Pptr, Pcol, Pval (prolongator n x nc, fp32 values)."""
        out = []
        for l in range(len(self.amg_levels())):
            sz = np.zeros(4, dtype=np.int64)
            self._check(self.lib.ddps_amg_get_level(self.h, l, _lib.iptr(sz), None, None, None, None, None))
            n, nnz, nc, pnnz = (int(v) for v in sz)
            lev = dict(n=n, nnz=nnz, nc=nc, rowptr=np.zeros(n + 1, np.int64), col=np.zeros(nnz, np.int64))
            if nc:
                lev.update(Pptr=np.zeros(n + 1, np.int64), Pcol=np.zeros(pnnz, np.int64), Pval=np.zeros(pnnz))
                self._check(self.lib.ddps_amg_get_level(self.h, l, _lib.iptr(sz), _lib.iptr(lev["rowptr"]), _lib.iptr(lev["col"]),
                                                        _lib.iptr(lev["Pptr"]), _lib.iptr(lev["Pcol"]), _lib.dptr(lev["Pval"])))
            else:
                self._check(self.lib.ddps_amg_get_level(self.h, l, _lib.iptr(sz), _lib.iptr(lev["rowptr"]), _lib.iptr(lev["col"]),
                                                        None, None, None))
            out.append(lev)
        return out

    def amg_numeric(self, which=MAT_OP):
        self._check(self.lib.ddps_amg_numeric(self.h, int(which)))

    def amg_get_values(self, level, nnz):
        """CSR-ordered values (``nnz`` of them) and inverse diagonal of multigrid level ``level`` >= 1."""
        n = self.amg_levels()[level]
        val, dinv = np.zeros(int(nnz)), np.zeros(n)
        self._check(self.lib.ddps_amg_get_values(self.h, int(level), _lib.dptr(val), _lib.dptr(dinv)))
        return val, dinv

    def amg_apply(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64)
        z = np.zeros_like(r)
        self._check(self.lib.ddps_amg_apply(self.h, _lib.dptr(r), _lib.dptr(z)))
        return z

    def time_kernel(self, kernel, warm=3, reps=20):
        ms, by = C.c_double(0), C.c_double(0)
        self._check(self.lib.ddps_time_kernel(self.h, int(kernel), int(warm), int(reps), C.byref(ms), C.byref(by)))
        return ms.value, by.value


# --------------------------------------------------------------------------------------------
# Reference-shaped entry points
# --------------------------------------------------------------------------------------------
def solve_ddpe(mesh, rec_mode, Vbddata, ubddata, vbddata, lam, delta, tau_p, tau_n, mu_p, mu_n, C_, tol=1e-9,
               session=None, return_stats=False, verbose=True):
    """Drop-in for ``solve_ddpe`` (src:630-658): returns ``(V, u, v)``.  Prints the same
    ``iteration k, tolerance: c`` lines (src:655) unless ``verbose=False``."""
    if str(rec_mode).lstrip(":") not in ("zero", "shockley"):
        raise ValueError("rec_mode must be :shockley or :zero (src:516-524)")
    own = session is None
    s = session or Session()
    try:
        if s.mesh is not mesh:
            s.set_mesh(mesh)
        s.setup_ddpe(Vbddata, ubddata, vbddata, lam, mu_p, mu_n, C_)
        s.set_opts(verbose=int(bool(verbose)))
        V, u, v, st = s.solve_ddpe(str(rec_mode).lstrip(":"), delta, tau_p, tau_n, tol)
        st["control_history"] = s.history(0)
        st["newton_per_vsolve"] = s.history(1)
        st["pcg_per_solve"] = s.history(2)
    finally:
        if own:
            s.close()
    return (V, u, v, st) if return_stats else (V, u, v)


def solve_semlin_poisson(mesh, A, bddata, f, g, tol=1e-14, session=None, return_stats=False, verbose=True):
    """Drop-in for ``solve_semlin_poisson`` (src:895-1097): returns the nodal solution.  ``f`` must
    be a registered functor; ``g`` is accepted for signature compatibility and must be
    ``f.derivative()`` (it is evaluated on the device from the functor's parameters)."""
    own = session is None
    s = session or Session()
    try:
        if s.mesh is not mesh:
            s.set_mesh(mesh)
        s.setup_semlin(A, bddata, f, g)
        s.set_opts(verbose=int(bool(verbose)))
        V, st = s.solve_semlin(tol)
        st["residual_history"] = s.history(0)
        st["pcg_per_solve"] = s.history(2)
    finally:
        if own:
            s.close()
    return (V, st) if return_stats else V
