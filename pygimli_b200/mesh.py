"""Flat-array mesh model consumed by the B200 ERT path.

The reference works on a pointer graph (``GIMLI::Mesh``: ``vector<Node*>``,
``vector<Cell*>``, ``vector<Boundary*>``; core/src/mesh.h:804-807).  The GPU path
needs structure-of-arrays; :class:`MeshArrays` is that model and the only mesh
type the kernels see.  ``from_pg_mesh`` converts a ``pg.Mesh`` when pygimli is
importable (drop-in use); the generators below build the synthetic tensor-grid
meshes named in SURVEY.md §8(d) (no Triangle/TetGen in this image).

Conventions mirrored from the reference:
  * node markers: -99 electrode, -999 reference electrode, -1000 calibration
    (core/src/bert/bert.h:30-32)
  * boundary markers: -1 homogeneous Neumann (surface), -2 mixed, -3 homogeneous
    Dirichlet (core/src/gimli.h:234-240)
  * 2D meshes live in the x-y plane, y is depth (dcfemmodelling.cpp:859-867)
  * P2 local numbering: Tri6 mid-nodes (0-1),(1-2),(2-0); Tet10 uses the
    Zienkiewicz order (0-1),(0-2),(0-3),(1-2),(2-3),(3-1)
    (core/src/meshentities.h:907-912)
  * H2 children inherit the parent's cell marker (core/src/mesh.cpp createRefined_)
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np

MARKER_NODE_ELECTRODE = -99
MARKER_NODE_REFERENCE = -999
MARKER_NODE_CALIBRATION = -1000
MARKER_BOUND_NEUMANN = -1
MARKER_BOUND_MIXED = -2
MARKER_BOUND_DIRICHLET = -3

TRI6_EDGES = ((0, 1), (1, 2), (2, 0))
TET10_EDGES = ((0, 1), (0, 2), (0, 3), (1, 2), (2, 3), (3, 1))


@dataclass
class MeshArrays:
    dim: int
    pos: np.ndarray            # (N, 3) float64
    node_marker: np.ndarray    # (N,)  int32
    cells: np.ndarray          # (C, nloc) int32, nloc in {3, 6, 4, 10}
    cell_marker: np.ndarray    # (C,)  int32
    bounds: np.ndarray         # (B, nlocb) int32 -- marked outer faces only
    bound_marker: np.ndarray   # (B,)  int32
    _cache: dict = field(default_factory=dict, repr=False)

    def __post_init__(self):
        self.pos = np.ascontiguousarray(self.pos, dtype=np.float64).reshape(-1, 3)
        self.node_marker = np.ascontiguousarray(self.node_marker, dtype=np.int32)
        self.cells = np.ascontiguousarray(self.cells, dtype=np.int32)
        self.cell_marker = np.ascontiguousarray(self.cell_marker, dtype=np.int32)
        self.bounds = np.ascontiguousarray(self.bounds, dtype=np.int32)
        self.bound_marker = np.ascontiguousarray(self.bound_marker, dtype=np.int32)
        if self.cells.ndim != 2 or self.cells.shape[1] not in (3, 4, 6, 10):
            raise ValueError("cells must be (C, nloc) with nloc in {3,6,4,10}")
        if (self.dim, self.cells.shape[1]) not in ((2, 3), (2, 6), (3, 4), (3, 10)):
            raise ValueError(f"unsupported cell type: dim={self.dim} nloc={self.cells.shape[1]}")

    # sizes ---------------------------------------------------------------
    @property
    def node_count(self) -> int:
        return self.pos.shape[0]

    @property
    def cell_count(self) -> int:
        return self.cells.shape[0]

    @property
    def nloc(self) -> int:
        return self.cells.shape[1]

    @property
    def nvert(self) -> int:
        """corner nodes per cell (3 triangles, 4 tetrahedra)"""
        return self.dim + 1

    @property
    def order(self) -> int:
        return 1 if self.nloc == self.nvert else 2

    # derived geometry ----------------------------------------------------
    def cell_sizes(self) -> np.ndarray:
        """area / volume of the straight-sided simplex (shape.cpp domainSize)."""
        p = self.pos[self.cells[:, : self.nvert]]
        if self.dim == 2:
            e1 = p[:, 1] - p[:, 0]
            e2 = p[:, 2] - p[:, 0]
            return 0.5 * np.abs(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0])
        e1 = p[:, 1] - p[:, 0]
        e2 = p[:, 2] - p[:, 0]
        e3 = p[:, 3] - p[:, 0]
        return np.abs(np.einsum("ij,ij->i", e1, np.cross(e2, e3))) / 6.0

    def bound_owner(self) -> np.ndarray:
        """owner cell of every listed boundary face (left cell of the reference,
        dcfemmodelling.cpp:256-258).  Found by matching corner-node sets."""
        if "bound_owner" in self._cache:
            return self._cache["bound_owner"]
        nvb = self.dim  # corner nodes of a face
        N = self.node_count
        if self.dim == 2:
            loc = ((0, 1), (1, 2), (2, 0))
        else:
            loc = ((0, 1, 2), (0, 1, 3), (1, 2, 3), (2, 0, 3))
        keys = []
        for lf in loc:
            f = np.sort(self.cells[:, list(lf)].astype(np.int64), axis=1)
            k = f[:, 0]
            for j in range(1, nvb):
                k = k * N + f[:, j]
            keys.append(k)
        keys = np.concatenate(keys)
        owner = np.tile(np.arange(self.cell_count, dtype=np.int64), len(loc))
        order = np.argsort(keys, kind="stable")
        keys, owner = keys[order], owner[order]
        fb = np.sort(self.bounds[:, :nvb].astype(np.int64), axis=1)
        kb = fb[:, 0]
        for j in range(1, nvb):
            kb = kb * N + fb[:, j]
        idx = np.searchsorted(keys, kb)
        if np.any(idx >= keys.size) or np.any(keys[np.minimum(idx, keys.size - 1)] != kb):
            raise ValueError("boundary face without an adjacent cell")
        res = owner[idx].astype(np.int32)
        self._cache["bound_owner"] = res
        return res


# ---------------------------------------------------------------------------
# axis helpers
# ---------------------------------------------------------------------------
def graded_axis(lo: float, hi: float, h: float, growth: float = 1.3, far: float = 0.0,
                both: bool = True) -> np.ndarray:
    """Uniform spacing ``h`` on [lo, hi], then geometric growth out to ``far``
    beyond each end (``both``) or beyond ``hi`` only."""
    n = int(round((hi - lo) / h))
    core = lo + h * np.arange(n + 1)
    out = [core]
    if far > 0:
        steps, d, tot = [], h, 0.0
        while tot < far:
            d *= growth
            tot += d
            steps.append(tot)
        steps = np.asarray(steps)
        out = ([lo - steps[::-1]] if both else []) + [core, hi + steps]
    return np.concatenate(out)


# ---------------------------------------------------------------------------
# 2-D: triangulated tensor grid
# ---------------------------------------------------------------------------
def grid_mesh_2d(x: np.ndarray, y: np.ndarray, para_box=None) -> MeshArrays:
    """Triangulate the tensor grid x (ascending) times y (descending from the surface y[0]).

    Each quad is split along alternating diagonals.  Boundary markers: top row -1
    (Neumann), the rest -2 (mixed).  Cell marker: consecutive index of the parent quad
    inside ``para_box=(xmin, xmax, ymin)`` (both triangles share it), -1 outside
    (background cells filled by prolongation, modellingbase.cpp:441-457).  Without
    ``para_box`` every quad is a model cell.
    """
    x = np.asarray(x, float)
    y = np.asarray(y, float)
    nx, ny = x.size, y.size
    X, Y = np.meshgrid(x, y)  # (ny, nx)
    pos = np.zeros((nx * ny, 3))
    pos[:, 0] = X.ravel()
    pos[:, 1] = Y.ravel()
    nid = lambda i, j: j * nx + i  # noqa: E731
    I, J = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1))
    I, J = I.ravel(), J.ravel()
    a, b, c, d = nid(I, J), nid(I + 1, J), nid(I + 1, J + 1), nid(I, J + 1)
    alt = ((I + J) % 2) == 0
    t1 = np.where(alt[:, None], np.stack([a, b, c], 1), np.stack([a, b, d], 1))
    t2 = np.where(alt[:, None], np.stack([a, c, d], 1), np.stack([b, c, d], 1))
    cells = np.empty((2 * I.size, 3), np.int32)
    cells[0::2] = t1
    cells[1::2] = t2
    # counter-clockwise orientation irrespective of the y direction
    p = pos[cells]
    det = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    flip = det < 0
    cells[flip] = cells[flip][:, [0, 2, 1]]

    xc = 0.5 * (x[I] + x[I + 1])
    yc = 0.5 * (y[J] + y[J + 1])
    if para_box is None:
        inside = np.ones(I.size, bool)
    else:
        xmin, xmax, ymin = para_box
        inside = (xc > xmin) & (xc < xmax) & (yc > ymin)
    qm = np.full(I.size, -1, np.int32)
    qm[inside] = np.arange(int(inside.sum()), dtype=np.int32)
    cell_marker = np.repeat(qm, 2)

    top = np.stack([nid(np.arange(nx - 1), 0), nid(np.arange(1, nx), 0)], 1)
    bot = np.stack([nid(np.arange(nx - 1), ny - 1), nid(np.arange(1, nx), ny - 1)], 1)
    lef = np.stack([nid(0, np.arange(ny - 1)), nid(0, np.arange(1, ny))], 1)
    rig = np.stack([nid(nx - 1, np.arange(ny - 1)), nid(nx - 1, np.arange(1, ny))], 1)
    bounds = np.concatenate([top, bot, lef, rig]).astype(np.int32)
    bmark = np.concatenate([np.full(len(top), MARKER_BOUND_NEUMANN), np.full(len(bot) + len(lef) + len(rig), MARKER_BOUND_MIXED)]).astype(np.int32)
    return MeshArrays(2, pos, np.zeros(nx * ny, np.int32), cells, cell_marker, bounds, bmark)


# ---------------------------------------------------------------------------
# 3-D: Kuhn-split tensor grid
# ---------------------------------------------------------------------------
_KUHN = np.array([[0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6], [0, 7, 4, 6], [0, 4, 5, 6], [0, 5, 1, 6]])


def grid_mesh_3d(x: np.ndarray, y: np.ndarray, z: np.ndarray, para_box=None, marker_per="cube") -> MeshArrays:
    """Kuhn 6-tetrahedra split of the tensor grid x*y*z (z descending from the surface z[0]).

    Boundary markers: top faces -1, all others -2.  Cell markers: ``marker_per='cube'`` gives
    the six tetrahedra of a hexahedron one model index (H2-like sharing), ``'cell'`` numbers
    every tetrahedron; ``para_box=(xmin,xmax,ymin,ymax,zmin)`` restricts the model region, the
    rest is background (-1).
    """
    x, y, z = (np.asarray(v, float) for v in (x, y, z))
    nx, ny, nz = x.size, y.size, z.size
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    pos = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
    nid = lambda i, j, k: (k * ny + j) * nx + i  # noqa: E731
    K, J, I = np.meshgrid(np.arange(nz - 1), np.arange(ny - 1), np.arange(nx - 1), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    v = np.stack([nid(I, J, K), nid(I + 1, J, K), nid(I + 1, J + 1, K), nid(I, J + 1, K),
                  nid(I, J, K + 1), nid(I + 1, J, K + 1), nid(I + 1, J + 1, K + 1), nid(I, J + 1, K + 1)], 1)
    cells = v[:, _KUHN].reshape(-1, 4).astype(np.int32)
    # positive orientation
    p = pos[cells]
    det = np.einsum("ij,ij->i", p[:, 1] - p[:, 0], np.cross(p[:, 2] - p[:, 0], p[:, 3] - p[:, 0]))
    flip = det < 0
    cells[flip] = cells[flip][:, [0, 2, 1, 3]]

    xc, yc, zc = 0.5 * (x[I] + x[I + 1]), 0.5 * (y[J] + y[J + 1]), 0.5 * (z[K] + z[K + 1])
    if para_box is None:
        inside = np.ones(I.size, bool)
    else:
        xmin, xmax, ymin, ymax, zmin = para_box
        inside = (xc > xmin) & (xc < xmax) & (yc > ymin) & (yc < ymax) & (zc > zmin)
    if marker_per == "cube":
        qm = np.full(I.size, -1, np.int32)
        qm[inside] = np.arange(int(inside.sum()), dtype=np.int32)
        cell_marker = np.repeat(qm, 6)
    else:
        ins = np.repeat(inside, 6)
        cell_marker = np.full(ins.size, -1, np.int32)
        cell_marker[ins] = np.arange(int(ins.sum()), dtype=np.int32)

    # outer faces: every tetra face on a grid plane. Collect per plane from the cube faces.
    def quad_faces(q):  # q: (n,4) corner ids of planar quads -> faces are found from the tets below
        return q

    faces, marks = [], []
    tri_loc = np.array([[0, 1, 2], [0, 1, 3], [1, 2, 3], [2, 0, 3]])
    f = cells[:, tri_loc].reshape(-1, 3)
    pf = pos[f]
    for axis, val, mark in ((2, z[0], MARKER_BOUND_NEUMANN), (2, z[-1], MARKER_BOUND_MIXED),
                            (0, x[0], MARKER_BOUND_MIXED), (0, x[-1], MARKER_BOUND_MIXED),
                            (1, y[0], MARKER_BOUND_MIXED), (1, y[-1], MARKER_BOUND_MIXED)):
        on = np.all(pf[:, :, axis] == val, axis=1)
        faces.append(f[on])
        marks.append(np.full(int(on.sum()), mark, np.int32))
    bounds = np.concatenate(faces).astype(np.int32)
    bmark = np.concatenate(marks)
    return MeshArrays(3, pos, np.zeros(pos.shape[0], np.int32), cells, cell_marker, bounds, bmark)


# ---------------------------------------------------------------------------
# refinement
# ---------------------------------------------------------------------------
def _edge_midnodes(mesh: MeshArrays, edges_loc):
    """unique mid-edge nodes in order of first appearance (cell order, local edge order)."""
    N = mesh.node_count
    c = mesh.cells[:, : mesh.nvert].astype(np.int64)
    e = np.stack([np.stack([c[:, a], c[:, b]], 1) for a, b in edges_loc], 1)  # (C, ne, 2)
    es = np.sort(e, axis=2).reshape(-1, 2)
    key = es[:, 0] * N + es[:, 1]
    uk, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # unique ids sorted by first appearance
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    mid_id = (N + rank[inv]).reshape(c.shape[0], len(edges_loc))
    ue = np.stack([uk // N, uk % N], 1)[order]
    mid_pos = 0.5 * (mesh.pos[ue[:, 0]] + mesh.pos[ue[:, 1]])
    return mid_id, mid_pos, uk[order]


def _lookup_mid(keys_sorted_by_rank, N, n0, n1, base):
    k = np.minimum(n0, n1).astype(np.int64) * N + np.maximum(n0, n1).astype(np.int64)
    srt = np.argsort(keys_sorted_by_rank)
    idx = np.searchsorted(keys_sorted_by_rank[srt], k)
    return base + srt[idx]


def create_p2(mesh: MeshArrays) -> MeshArrays:
    """Quadratic (P2) copy of a P1 mesh (what Mesh::createP2 does, mesh.cpp:1315)."""
    if mesh.order != 1:
        raise ValueError("create_p2 needs a P1 mesh")
    edges = TRI6_EDGES if mesh.dim == 2 else TET10_EDGES
    mid_id, mid_pos, keys = _edge_midnodes(mesh, edges)
    N = mesh.node_count
    pos = np.concatenate([mesh.pos, mid_pos])
    nm = np.concatenate([mesh.node_marker, np.zeros(mid_pos.shape[0], np.int32)])
    cells = np.concatenate([mesh.cells, mid_id.astype(np.int32)], 1)
    b = mesh.bounds
    if mesh.dim == 2:
        bm = _lookup_mid(keys, N, b[:, 0], b[:, 1], N)
        bounds = np.concatenate([b, bm[:, None].astype(np.int32)], 1)
    else:
        m01 = _lookup_mid(keys, N, b[:, 0], b[:, 1], N)
        m12 = _lookup_mid(keys, N, b[:, 1], b[:, 2], N)
        m20 = _lookup_mid(keys, N, b[:, 2], b[:, 0], N)
        bounds = np.concatenate([b, np.stack([m01, m12, m20], 1).astype(np.int32)], 1)
    return MeshArrays(mesh.dim, pos, nm, cells, mesh.cell_marker.copy(), bounds, mesh.bound_marker.copy())


def create_h2(mesh: MeshArrays) -> MeshArrays:
    """Uniform h-refinement (tri -> 4, tet -> 8 children; Mesh::createH2, mesh.cpp:1278).
    Children inherit the parent's marker, so model columns are shared."""
    if mesh.order != 1:
        raise ValueError("create_h2 needs a P1 mesh")
    edges = TRI6_EDGES if mesh.dim == 2 else TET10_EDGES
    mid_id, mid_pos, keys = _edge_midnodes(mesh, edges)
    N = mesh.node_count
    pos = np.concatenate([mesh.pos, mid_pos])
    nm = np.concatenate([mesh.node_marker, np.zeros(mid_pos.shape[0], np.int32)])
    c = mesh.cells.astype(np.int64)
    m = mid_id
    if mesh.dim == 2:
        ch = [np.stack([c[:, 0], m[:, 0], m[:, 2]], 1), np.stack([m[:, 0], c[:, 1], m[:, 1]], 1),
              np.stack([m[:, 2], m[:, 1], c[:, 2]], 1), np.stack([m[:, 0], m[:, 1], m[:, 2]], 1)]
        nch = 4
    else:
        # m: 0:(0-1) 1:(0-2) 2:(0-3) 3:(1-2) 4:(2-3) 5:(3-1)
        ch = [np.stack([c[:, 0], m[:, 0], m[:, 1], m[:, 2]], 1), np.stack([m[:, 0], c[:, 1], m[:, 3], m[:, 5]], 1),
              np.stack([m[:, 1], m[:, 3], c[:, 2], m[:, 4]], 1), np.stack([m[:, 2], m[:, 5], m[:, 4], c[:, 3]], 1),
              # inner octahedron split along the (0-2)-(3-1) diagonal  m1-m5
              np.stack([m[:, 0], m[:, 3], m[:, 1], m[:, 5]], 1), np.stack([m[:, 0], m[:, 1], m[:, 2], m[:, 5]], 1),
              np.stack([m[:, 1], m[:, 3], m[:, 4], m[:, 5]], 1), np.stack([m[:, 1], m[:, 4], m[:, 2], m[:, 5]], 1)]
        nch = 8
    cells = np.stack(ch, 1).reshape(-1, mesh.nvert).astype(np.int32)
    p = pos[cells]
    if mesh.dim == 2:
        det = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
        flip = det < 0
        cells[flip] = cells[flip][:, [0, 2, 1]]
    else:
        det = np.einsum("ij,ij->i", p[:, 1] - p[:, 0], np.cross(p[:, 2] - p[:, 0], p[:, 3] - p[:, 0]))
        flip = det < 0
        cells[flip] = cells[flip][:, [0, 2, 1, 3]]
    cm = np.repeat(mesh.cell_marker, nch)
    b = mesh.bounds
    if mesh.dim == 2:
        bm = _lookup_mid(keys, N, b[:, 0], b[:, 1], N).astype(np.int32)
        bounds = np.stack([np.stack([b[:, 0], bm], 1), np.stack([bm, b[:, 1]], 1)], 1).reshape(-1, 2)
        bmark = np.repeat(mesh.bound_marker, 2)
    else:
        m01 = _lookup_mid(keys, N, b[:, 0], b[:, 1], N).astype(np.int32)
        m12 = _lookup_mid(keys, N, b[:, 1], b[:, 2], N).astype(np.int32)
        m20 = _lookup_mid(keys, N, b[:, 2], b[:, 0], N).astype(np.int32)
        bounds = np.stack([np.stack([b[:, 0], m01, m20], 1), np.stack([m01, b[:, 1], m12], 1),
                           np.stack([m20, m12, b[:, 2]], 1), np.stack([m01, m12, m20], 1)], 1).reshape(-1, 3)
        bmark = np.repeat(mesh.bound_marker, 4)
    return MeshArrays(mesh.dim, pos, nm, cells, cm, bounds, bmark)


def mark_electrode_nodes(mesh: MeshArrays, sensors: np.ndarray, tol: float = 1e-6) -> np.ndarray:
    """Set marker -99 on the nodes coinciding with the sensor positions; returns node ids (-1 if none)."""
    sensors = np.asarray(sensors, float).reshape(-1, 3)
    ids = np.full(sensors.shape[0], -1, np.int64)
    for i, s in enumerate(sensors):
        d2 = np.sum((mesh.pos - s) ** 2, axis=1)
        j = int(np.argmin(d2))
        if d2[j] < tol * tol:
            ids[i] = j
            mesh.node_marker[j] = MARKER_NODE_ELECTRODE
    return ids


def from_pg_mesh(pgmesh) -> MeshArrays:
    """Convert a ``pg.Mesh`` (only when pygimli is importable).  Uses the vectorised
    accessors of the binding; boundaries with marker 0 are dropped."""
    import numpy as _np
    pos = _np.asarray(pgmesh.positions().array() if hasattr(pgmesh.positions(), "array") else pgmesh.positions())
    nm = _np.asarray(pgmesh.nodeMarkers())
    cells = _np.asarray([[n.id() for n in c.nodes()] for c in pgmesh.cells()], dtype=_np.int32)
    cm = _np.asarray(pgmesh.cellMarkers())
    bl = [(b, b.marker()) for b in pgmesh.boundaries() if b.marker() != 0]
    bounds = _np.asarray([[n.id() for n in b.nodes()] for b, _ in bl], dtype=_np.int32)
    bm = _np.asarray([m for _, m in bl], dtype=_np.int32)
    return MeshArrays(int(pgmesh.dim()), pos, nm, cells, cm, bounds, bm)
