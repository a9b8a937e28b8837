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
