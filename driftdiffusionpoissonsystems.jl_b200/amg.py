"""Host-side view of the smoothed-aggregation hierarchy through the library's host-only
entry points (``ddps_amg_host_*``, no GPU needed): used by the tests to check the frozen prolongators and
coarse patterns against scipy, and handy for inspecting coarsening on a user's mesh."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def host_hierarchy(rowptr, col, val, fixed, nthreads=0, coarse_max=0):
    """Build the hierarchy of the CSR matrix (0-based, sorted columns); ``fixed`` flags Dirichlet rows.
    Returns a list of per-level dicts: n, rowptr, col, val (base operator of the level) and, except on
    the coarsest level, nc, Pptr, Pcol, Pval (prolongator n x nc) and agg (aggregate id per row, -1 none)."""
    lib = _lib.load()
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    val = np.ascontiguousarray(val, dtype=np.float64)
    fixed = np.ascontiguousarray(fixed, dtype=np.uint8)
    n = rowptr.size - 1
    h = C.c_void_p()
    rc = lib.ddps_amg_host_create(C.byref(h), n, _lib.iptr(rowptr), _lib.iptr(col), _lib.dptr(val),
                                  fixed.ctypes.data_as(_lib.c_uint8_p), int(nthreads), int(coarse_max))
    if rc != 0:
        raise RuntimeError(f"ddps_amg_host_create failed ({rc})")
    try:
        levels = []
        for l in range(lib.ddps_amg_host_nlevels(h)):
            sz = np.zeros(4, dtype=np.int64)
            lib.ddps_amg_host_level_sizes(h, l, _lib.iptr(sz))
            nl, nnz, nc, pnnz = (int(v) for v in sz)
            lev = dict(n=nl, nc=nc, rowptr=np.zeros(nl + 1, np.int64), col=np.zeros(nnz, np.int64), val=np.zeros(nnz))
            if nc:
                lev.update(Pptr=np.zeros(nl + 1, np.int64), Pcol=np.zeros(pnnz, np.int64), Pval=np.zeros(pnnz),
                           agg=np.zeros(nl, np.int64))
                lib.ddps_amg_host_level_get(h, l, _lib.iptr(lev["rowptr"]), _lib.iptr(lev["col"]), _lib.dptr(lev["val"]),
                                            _lib.iptr(lev["Pptr"]), _lib.iptr(lev["Pcol"]), _lib.dptr(lev["Pval"]),
                                            _lib.iptr(lev["agg"]))
            else:
                lib.ddps_amg_host_level_get(h, l, _lib.iptr(lev["rowptr"]), _lib.iptr(lev["col"]), _lib.dptr(lev["val"]),
                                            None, None, None, None)
            levels.append(lev)
        return levels
    finally:
        lib.ddps_amg_host_free(h)
