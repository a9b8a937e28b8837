"""Mesh container, gmsh reader and .geo writers: the reference's L0 layer, kept on the host.

``/root/reference/src/DriftDiffusionPoissonSystems.jl`` is cited as ``src:N``.  These are host I/O
routines (one-time, SURVEY.md section 2 marks them out of scope for acceleration); they exist so
that user code written against the reference keeps working.
"""
from __future__ import annotations

import dataclasses
import warnings

import numpy as np


@dataclasses.dataclass
class Mesh:
    """Same fields and layout as the reference's ``Mesh`` type (src:9-18), with the Julia
    column-major arrays stored as (rows, count) numpy arrays:
    nodes 2 x nq float64; edges 4 x ne int64 [n1;n2;phys;geom]; elements 5 x nme int64
    [n1;n2;n3;phys;geom]; node ids 1-based."""

    nodes: np.ndarray
    edges: np.ndarray
    elements: np.ndarray

    @property
    def nq(self):
        return self.nodes.shape[1]

    @property
    def nme(self):
        return self.elements.shape[1]


def read_mesh(path) -> Mesh:
    """gmsh MSH 2.2 ASCII reader with the acceptance rules of src:95-154: format 2.2, planar
    nodes, only 2-node lines (type 1) and 3-node triangles (type 2) with exactly two tags; anything
    after ``$EndElements`` is ignored with a warning (src:149-151)."""
    with open(path, "r") as fh:
        tok = [ln.split() for ln in fh]
    pos = 0

    def expect(word):
        nonlocal pos
        if pos >= len(tok) or tok[pos] != [word]:
            raise AssertionError(f"{path}: expected {word} at line {pos + 1}")
        pos += 1

    expect("$MeshFormat")
    if float(tok[pos][0]) != 2.2:
        raise AssertionError("only MSH format 2.2 is supported (src:103)")
    pos += 1
    expect("$EndMeshFormat")
    if tok[pos] == ["$PhysicalNames"]:
        pos += int(tok[pos + 1][0]) + 3
    expect("$Nodes")
    nq = int(tok[pos][0])
    block = np.array(tok[pos + 1:pos + 1 + nq], dtype=np.float64)
    if np.any(block[:, 3] != 0.0):
        raise AssertionError("mesh is not planar (src:119)")
    nodes = np.ascontiguousarray(block[:, 1:3].T)
    pos += nq + 1
    expect("$EndNodes")
    expect("$Elements")
    nall = int(tok[pos][0])
    edges, tris = [], []
    for rec in tok[pos + 1:pos + 1 + nall]:
        kind = int(rec[1])
        if kind not in (1, 2):
            raise ValueError("Unknown type of elements used! Please only use edges and triangular elements!")
        if int(rec[2]) != 2:
            raise AssertionError("elements must carry exactly two tags (src:135,137)")
        if kind == 1:
            edges.append((rec[5], rec[6], rec[3], rec[4]))
        else:
            tris.append((rec[5], rec[6], rec[7], rec[3], rec[4]))
    pos += nall + 1
    expect("$EndElements")
    if pos < len(tok):
        warnings.warn("There are lines in the mesh file that are not needed and cannot be taken into account.")
    e = np.array(edges, dtype=np.int64).reshape(-1, 4).T
    t = np.array(tris, dtype=np.int64).reshape(-1, 5).T
    return Mesh(nodes, np.ascontiguousarray(e), np.ascontiguousarray(t))


def _geo_header(h):
    return [f"h = {h};", " "]


def construct_mesh4(Endpoints, h, meshname):
    """Write ``meshname.geo`` for a rectangle with 4 boundary lines (src:63-88).  Run
    ``gmsh -2 meshname.geo`` yourself to obtain the ``.msh``."""
    E = np.asarray(Endpoints, dtype=float)
    out = _geo_header(h)
    for k in range(4):
        out.append(f"Point({k + 1})={{{E[k, 0]}, {E[k, 1]}, 0, h}};")
    for k in range(4):
        out.append(f"Line({k + 1}) = {{{k + 1}, {(k + 1) % 4 + 1}}};")
    out += ["Line Loop(1) = {1,2,3,4};", "Plane Surface(2) ={1};", "Physical Surface(2) = {2};",
            "Physical Line(1) = {1, 2, 3, 4};", "Coherence;"]
    with open(f"{meshname}.geo", "w") as fh:
        fh.write("\n".join(out))


def construct_mesh8(Endpoints, Dirpoints, h, meshname):
    """Write ``meshname.geo`` for a rectangle whose left and right sides are split into three
    segments each (8 boundary lines, src:24-57); ``Dirpoints`` gives the y-extent of the middle
    (contact) segments: row 1 left side (segment 7), row 2 right side (segment 3)."""
    E = np.asarray(Endpoints, dtype=float)
    D = np.asarray(Dirpoints, dtype=float)
    pts = [(E[0, 0], E[0, 1]), (E[1, 0], E[1, 1]), (E[1, 0], D[1, 0]), (E[1, 0], D[1, 1]),
           (E[2, 0], E[2, 1]), (E[3, 0], E[3, 1]), (E[3, 0], D[0, 1]), (E[3, 0], D[0, 0])]
    out = _geo_header(h)
    for k, (x, y) in enumerate(pts):
        out.append(f"Point({k + 1})={{{x}, {y}, 0, h}};")
    for k in range(8):
        out.append(f"Line({k + 1}) = {{{k + 1}, {(k + 1) % 8 + 1}}};")
    out += ["Line Loop(1) = {1,2,3,4,5,6,7,8};", "Plane Surface(2) ={1};", "Physical Surface(2) = {2};",
            "Physical Line(1) = {1, 2, 3, 4, 5, 6, 7, 8};", "Coherence;"]
    with open(f"{meshname}.geo", "w") as fh:
        fh.write("\n".join(out))


# README.md:60 and test/runtests.jl:10 call ``construct_mesh`` although the module only exports the
# 4/8 variants (src:6); alias the 8-segment writer so those scripts run.
construct_mesh = construct_mesh8


def structured_mesh(nx, ny, x0=0.0, x1=1.0, y0=0.0, y1=1.0) -> Mesh:
    """Synthetic structured triangulation in the reference's Mesh layout (benchmark configs 3-5,
    SURVEY.md section 8d): (nx+1)(ny+1) nodes, node (i,j) -> id j(nx+1)+i+1, every cell split along
    its (i,j)-(i+1,j+1) diagonal into two CCW triangles (n00,n10,n11),(n00,n11,n01); boundary edges
    tagged 1 bottom, 2 right, 3 top, 4 left."""
    xs = np.linspace(x0, x1, nx + 1)
    ys = np.linspace(y0, y1, ny + 1)
    # stored like the reference's Julia arrays (column-major 2 x nq / 5 x nme = x,y interleaved per node, the five
    # entries of a triangle adjacent): the (rows, count) views below are Fortran-ordered, and the C ABI takes them as is
    nodes = np.empty(((nx + 1) * (ny + 1), 2)).T
    nodes[0] = np.tile(xs, ny + 1)
    nodes[1] = np.repeat(ys, nx + 1)
    cell = (np.arange(ny, dtype=np.int64)[:, None] * (nx + 1) + np.arange(nx, dtype=np.int64)[None, :]).ravel() + 1
    el = np.empty((2 * nx * ny, 5), dtype=np.int64).T
    el[0, 0::2] = cell
    el[1, 0::2] = cell + 1
    el[2, 0::2] = cell + nx + 2
    el[0, 1::2] = cell
    el[1, 1::2] = cell + nx + 2
    el[2, 1::2] = cell + nx + 1
    el[3:] = 2
    i = np.arange(nx, dtype=np.int64)
    j = np.arange(ny, dtype=np.int64)
    one = lambda n, t: (np.full(n, 1, dtype=np.int64), np.full(n, t, dtype=np.int64))
    bottom = np.vstack([i + 1, i + 2, *one(nx, 1)])
    right = np.vstack([j * (nx + 1) + nx + 1, (j + 1) * (nx + 1) + nx + 1, *one(ny, 2)])
    top = np.vstack([ny * (nx + 1) + nx + 1 - i, ny * (nx + 1) + nx - i, *one(nx, 3)])
    left = np.vstack([(ny - j) * (nx + 1) + 1, (ny - j - 1) * (nx + 1) + 1, *one(ny, 4)])
    return Mesh(nodes, np.ascontiguousarray(np.hstack([bottom, right, top, left])), el)


# --------------------------------------------------------------------------------------------
# Per-rank parts of a structured mesh (multi-GPU runs where no process holds the global mesh;
# C ABI: ddps_mesh_set_local)
# --------------------------------------------------------------------------------------------
class _StructuredNodes:
    """Lazy stand-in for ``Mesh.nodes`` of a structured mesh: ``nodes[k, ids]`` (k = 0: x, 1: y; ids = 0-based node ids)
    computes the coordinates instead of storing 2 x nq doubles.  Same values as ``structured_mesh`` bit for bit."""

    def __init__(self, nx, ny, x0, x1, y0, y1):
        self.nx, self.xs, self.ys = nx, np.linspace(x0, x1, nx + 1), np.linspace(y0, y1, ny + 1)
        self.shape = (2, (nx + 1) * (ny + 1))

    def __getitem__(self, key):
        k, ids = key
        ids = np.asarray(ids, dtype=np.int64)
        return self.xs[ids % (self.nx + 1)] if k == 0 else self.ys[ids // (self.nx + 1)]


def structured_boundary(nx, ny, x0=0.0, x1=1.0, y0=0.0, y1=1.0) -> Mesh:
    """The boundary of ``structured_mesh``: its edges (tags 1 bottom, 2 right, 3 top, 4 left) and lazily computed node
    coordinates -- all that the host-side boundary preprocessing (``dirichlet_nodes``, src:162-218) reads."""
    i = np.arange(nx, dtype=np.int64)
    j = np.arange(ny, dtype=np.int64)
    one = lambda n, t: (np.full(n, 1, dtype=np.int64), np.full(n, t, dtype=np.int64))
    bottom = np.vstack([i + 1, i + 2, *one(nx, 1)])
    right = np.vstack([j * (nx + 1) + nx + 1, (j + 1) * (nx + 1) + nx + 1, *one(ny, 2)])
    top = np.vstack([ny * (nx + 1) + nx + 1 - i, ny * (nx + 1) + nx - i, *one(nx, 3)])
    left = np.vstack([(ny - j) * (nx + 1) + 1, (ny - j - 1) * (nx + 1) + 1, *one(ny, 4)])
    return Mesh(_StructuredNodes(nx, ny, x0, x1, y0, y1), np.ascontiguousarray(np.hstack([bottom, right, top, left])), None)


def structured_mesh_part(nx, ny, x0, x1, y0, y1, rank, nranks):
    """Rank ``rank``'s part of ``structured_mesh(nx, ny, ...)`` under a partition of the node COLUMNS into ``nranks``
    vertical strips, built directly (the global mesh is never formed), in the layout ``ddps_mesh_set_local`` takes:
    owned nodes in increasing global id, ghost nodes (the two neighbouring node columns) sorted by (owner, global id),
    every triangle touching an owned node in increasing global triangle id with LOCAL 1-based node ids."""
    ncol = nx + 1
    cut = [r * ncol // nranks for r in range(nranks + 1)]
    i0, i1 = cut[rank], cut[rank + 1]
    if i1 - i0 < 2 and nranks > 1:
        raise ValueError("strips narrower than two node columns are not supported")
    xs, ys = np.linspace(x0, x1, nx + 1), np.linspace(y0, y1, ny + 1)
    jj = np.arange(ny + 1, dtype=np.int64)
    own_i = np.arange(i0, i1, dtype=np.int64)
    own_glob = (jj[:, None] * ncol + own_i[None, :]).ravel()                 # increasing global id (j-major)
    ghost_cols, ghost_owner_of_col = [], []
    if i0 > 0:
        ghost_cols.append(i0 - 1); ghost_owner_of_col.append(rank - 1)
    if i1 < ncol:
        ghost_cols.append(i1); ghost_owner_of_col.append(rank + 1)
    n_own = own_glob.size
    ghost_glob = [jj * ncol + c for c in ghost_cols]                          # per owner: increasing global id
    node_glob = np.concatenate([own_glob] + ghost_glob) + 1
    ghost_owner = np.concatenate([np.full(ny + 1, o, dtype=np.int32) for o in ghost_owner_of_col]) if ghost_cols else np.zeros(0, np.int32)
    n_loc = node_glob.size
    nodes = np.empty((n_loc, 2))
    g0 = node_glob - 1
    nodes[:, 0] = xs[g0 % ncol]
    nodes[:, 1] = ys[g0 // ncol]
    # local id of a node (column i, row j): owned -> j*(i1-i0) + (i-i0); ghost column k -> n_own + k*(ny+1) + j
    w = i1 - i0

    def local_id(i, j):
        out = j * w + (i - i0)
        for k, c in enumerate(ghost_cols):
            out = np.where(i == c, n_own + k * (ny + 1) + j, out)
        return out + 1

    ci0, ci1 = max(i0 - 1, 0), min(i1, nx)                                    # cell columns [ci0, ci1) touch an owned node column
    ci = np.arange(ci0, ci1, dtype=np.int64)
    cj = np.arange(ny, dtype=np.int64)
    I = np.broadcast_to(ci[None, :], (ny, ci.size)).ravel()
    J = np.broadcast_to(cj[:, None], (ny, ci.size)).ravel()
    n00, n10, n11, n01 = local_id(I, J), local_id(I + 1, J), local_id(I + 1, J + 1), local_id(I, J + 1)
    nme_loc = 2 * I.size
    el = np.empty((nme_loc, 5), dtype=np.int64)
    el[0::2, 0], el[0::2, 1], el[0::2, 2] = n00, n10, n11
    el[1::2, 0], el[1::2, 1], el[1::2, 2] = n00, n11, n01
    el[:, 3:] = 2
    elem_glob = np.empty(nme_loc, dtype=np.int64)
    cell = J * nx + I
    elem_glob[0::2] = 2 * cell + 1
    elem_glob[1::2] = 2 * cell + 2
    return dict(nodes=nodes, n_own=int(n_own), n_loc=int(n_loc), node_glob=np.ascontiguousarray(node_glob, dtype=np.int64),
                ghost_owner=np.ascontiguousarray(ghost_owner, dtype=np.int32), elements=el, elem_glob=elem_glob,
                nq_global=(nx + 1) * (ny + 1), nme_global=2 * nx * ny, own_glob=own_glob,
                boundary=structured_boundary(nx, ny, x0, x1, y0, y1))
