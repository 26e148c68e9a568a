"""One-off host-side set-up for the B200 ERT path (geometry-only, runs once per mesh/scheme).

Everything here is *plan building*: it produces the flat arrays the CUDA kernels consume
(CSR pattern + scatter map + colours, boundary-coefficient tables, electrode tables,
wavenumber lists, prolongation levels, Jacobian column segments).  No resistivity-dependent
arithmetic happens on the host; the per-call work is all on the GPU.

Reference semantics restated (file:line under /root/reference/core/src):
  pattern            sparsematrix.h:966-1032  (union of all node pairs per cell, cols ascending)
  wavenumbers        bert/bertMisc.cpp:36-129, numericbase.cpp:51-150
  mixed BC           bert/dcfemmodelling.cpp:243-299, :430-506
  electrodes         bert/dcfemmodelling.cpp:790-1070, bert/electrode.cpp:102-287
  prolongation       modellingbase.cpp:401-497, mesh.cpp:2247-2316
  Jacobian columns   bert/bertJacobian.cpp:280-299
"""
from __future__ import annotations

import math
import numpy as np

from .mesh import (MeshArrays, MARKER_NODE_ELECTRODE, MARKER_NODE_REFERENCE, MARKER_NODE_CALIBRATION,
                   MARKER_BOUND_NEUMANN, MARKER_BOUND_MIXED, MARKER_BOUND_DIRICHLET)

TOLERANCE = 1e-12
MAX_DOUBLE = np.finfo(np.float64).max


# ---------------------------------------------------------------------------
# Bessel functions: Abramowitz & Stegun 9.8.1-9.8.8 polynomial approximations, the
# same formulas (and therefore the same ~1e-7 accuracy) the reference uses
# (numericbase.h:80-180).  Vectorised numpy; the CUDA twin lives in csrc/ert_device.cuh.
# ---------------------------------------------------------------------------
def bessel_i0(x):
    x = np.asarray(x, float)
    ax = np.abs(x)
    y = (x / 3.75) ** 2
    small = 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        yy = 3.75 / ax
        big = (np.exp(ax) / np.sqrt(ax)) * (0.39894228 + yy * (0.1328592e-1 + yy * (0.225319e-2 + yy * (-0.157565e-2 + yy * (
            0.916281e-2 + yy * (-0.2057706e-1 + yy * (0.2635537e-1 + yy * (-0.1647633e-1 + yy * 0.392377e-2))))))))
    return np.where(ax < 3.75, small, big)


def bessel_i1(x):
    x = np.asarray(x, float)
    ax = np.abs(x)
    y = (x / 3.75) ** 2
    small = ax * (0.5 + y * (0.87890594 + y * (0.51498869 + y * (0.15084934 + y * (0.2658733e-1 + y * (0.301532e-2 + y * 0.32411e-3))))))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        yy = 3.75 / ax
        r = 0.2282967e-1 + yy * (-0.2895312e-1 + yy * (0.1787654e-1 - yy * 0.420059e-2))
        r = 0.39894228 + yy * (-0.3988024e-1 + yy * (-0.362018e-2 + yy * (0.163801e-2 + yy * (-0.1031555e-1 + yy * r))))
        big = r * (np.exp(ax) / np.sqrt(ax))
    res = np.where(ax < 3.75, small, big)
    return np.where(x < 0.0, -res, res)


def bessel_k0(x):
    x = np.asarray(x, float)
    with np.errstate(divide="ignore", invalid="ignore"):
        y = x * x / 4.0
        small = (-np.log(x / 2.0) * bessel_i0(x)) + (-0.57721566 + y * (0.42278420 + y * (0.23069756 + y * (
            0.3488590e-1 + y * (0.262698e-2 + y * (0.10750e-3 + y * 0.74e-5))))))
        yy = 2.0 / x
        big = (np.exp(-x) / np.sqrt(x)) * (1.25331414 + yy * (-0.7832358e-1 + yy * (0.2189568e-1 + yy * (
            -0.1062446e-1 + yy * (0.587872e-2 + yy * (-0.251540e-2 + yy * 0.53208e-3))))))
    return np.where(x <= 2.0, small, big)


def bessel_k1(x):
    x = np.asarray(x, float)
    with np.errstate(divide="ignore", invalid="ignore"):
        y = x * x / 4.0
        small = (np.log(x / 2.0) * bessel_i1(x)) + (1.0 / x) * (1.0 + y * (0.15443144 + y * (-0.67278579 + y * (
            -0.18156897 + y * (-0.1919402e-1 + y * (-0.110404e-2 + y * (-0.4686e-4)))))))
        yy = 2.0 / x
        big = (np.exp(-x) / np.sqrt(x)) * (1.25331414 + yy * (0.23498619 + yy * (-0.3655620e-1 + yy * (
            0.1504268e-1 + yy * (-0.780353e-2 + yy * (0.325614e-2 + yy * (-0.68245e-3)))))))
    return np.where(x <= 2.0, small, big)


# ---------------------------------------------------------------------------
# Gauss rules exactly as the reference iterates them (loose Newton tolerances kept on
# purpose: the wavenumber weights must match the reference's, not the exact rules).
# ---------------------------------------------------------------------------
def gauss_legendre(x1: float, x2: float, n: int):
    """numericbase.cpp:102-150 (Newton tolerance 3e-6; weight from the pre-update derivative)."""
    x = np.zeros(n)
    w = np.zeros(n)
    eps = 3.0e-6
    m = (n + 1.0) / 2.0
    xm = 0.5 * (x2 + x1)
    xl = 0.5 * (x2 - x1)
    i = 1
    while i <= m:
        z = math.cos(math.pi * (i - 0.25) / (n + 0.5))
        z1 = z + 2.0 * eps
        pp = 0.0
        while abs(z - z1) > eps:
            p1, p2 = 1.0, 0.0
            for j in range(1, n + 1):
                p3 = p2
                p2 = p1
                p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / float(j)
            pp = float(n) * (z * p1 - p2) / (z * z - 1.0)
            z1 = z
            z = z1 - p1 / pp
        x[i - 1] = xm - xl * z
        x[n - i] = xm + xl * z
        w[i - 1] = 2.0 * xl / ((1.0 - z * z) * pp * pp)
        w[n - i] = w[i - 1]
        i += 1
    return x, w


def gauss_laguerre(n: int):
    """numericbase.cpp:51-100 (alpha = 0, tolerance 3e-11, at most 20 Newton steps)."""
    x = np.zeros(n)
    w = np.zeros(n)
    eps = 3.0e-11
    z = 0.0
    for i in range(1, n + 1):
        if i == 1:
            z = 3.0 / (1.0 + 2.4 * n)
        elif i == 2:
            z = z + 15.0 / (1.0 + 2.5 * n)
        else:
            ai = i - 2
            z = z + (1.0 + 2.55 * ai) / (1.9 * ai) * (z - x[ai - 1])
        pp = p2 = 0.0
        for _ in range(20):
            p1, p2 = 1.0, 0.0
            for j in range(1, n + 1):
                p3 = p2
                p2 = p1
                p1 = ((2.0 * j - 1 - z) * p2 - (j - 1) * p3) / j
            pp = n * (p1 - p2) / z
            z1 = z
            z = z1 - p1 / pp
            if abs(z - z1) <= eps:
                break
        x[i - 1] = z
        w[i - 1] = -1.0 / (pp * n * p2)
    return x, w


def init_kwave_list(dim: int, sensors: np.ndarray):
    """bertMisc.cpp:36-129: 3-D -> k={0}, w={1}; 2-D -> Legendre + Laguerre mix."""
    if dim == 3:
        return np.zeros(1), np.ones(1)
    s = np.asarray(sensors, float).reshape(-1, 3)
    ne = s.shape[0]
    if ne < 2:
        raise ValueError("need at least two sensors to initialise the wavenumber list")
    d = np.sqrt(((s[:, None, :] - s[None, :, :]) ** 2).sum(-1))
    iu = np.triu_indices(ne, 1)
    rmin = d[iu].min() / 2.0
    rmax = d[iu].max() * 2.0
    nleg = max(int(math.floor(6.0 * math.log10(rmax / rmin))), 4)
    nlag = 4
    return kwave_from_range(rmin, rmax, nleg, nlag)


def kwave_from_range(rmin, rmax, nleg, nlag):
    k0 = 1.0 / (2.0 * rmin)
    x, w = gauss_legendre(0.0, 1.0, nleg)
    kleg = k0 * x * x
    wleg = 2.0 * k0 * x * w / math.pi
    x, w = gauss_laguerre(nlag)
    klag = k0 * (x + 1.0)
    wlag = k0 * np.exp(x) * w / math.pi
    return np.concatenate([kleg, klag]), np.concatenate([wleg, wlag])


# ---------------------------------------------------------------------------
# CSR pattern + scatter map
# ---------------------------------------------------------------------------
def build_pattern(mesh: MeshArrays):
    """rowptr[N+1], colidx[nnz] (int32, columns ascending per row) and the per-cell scatter
    map pos[C, nloc*nloc] (CSR slot of local entry (i, j))."""
    N = mesh.node_count
    c = mesh.cells.astype(np.int64)
    nloc = c.shape[1]
    rows = np.repeat(c, nloc, axis=1)            # i-major: (i, j) -> row c[i]
    cols = np.tile(c, (1, nloc))                 #                   col c[j]
    keys = (rows * N + cols).ravel()
    ukeys = np.unique(keys)
    if ukeys.size >= 2 ** 31:
        raise OverflowError("pattern exceeds int32 index range")
    colidx = (ukeys % N).astype(np.int32)
    counts = np.bincount((ukeys // N), minlength=N)
    rowptr = np.zeros(N + 1, np.int32)
    np.cumsum(counts, out=rowptr[1:])
    pos = np.searchsorted(ukeys, keys).astype(np.int32).reshape(c.shape[0], nloc * nloc)
    return rowptr, colidx, pos


def csr_positions(rowptr, colidx, rows, cols):
    """CSR slot of every (row, col) pair (must exist)."""
    N = rowptr.size - 1
    rowof = np.repeat(np.arange(N, dtype=np.int64), np.diff(rowptr))
    keys = rowof * N + colidx
    q = np.asarray(rows, np.int64) * N + np.asarray(cols, np.int64)
    p = np.searchsorted(keys, q)
    if np.any(keys[np.minimum(p, keys.size - 1)] != q):
        raise KeyError("requested entry not in the sparsity pattern")
    return p.astype(np.int32)


def color_cells_numpy(cells: np.ndarray, n_nodes: int, max_colors: int = 256):
    """Greedy conflict colouring (two cells conflict when they share a node), vectorised as
    repeated maximal-independent-set extraction.  The C++ helper pgb200_color_cells is the
    fast path; this is the portable twin used when the library is not built."""
    C, nloc = cells.shape
    rng = np.random.default_rng(7)
    prio = rng.permutation(C).astype(np.int64) + 1
    color = np.full(C, -1, np.int32)
    remaining = np.arange(C)
    col = 0
    while remaining.size:
        if col >= max_colors:
            raise RuntimeError("colouring needs too many colours")
        cand = remaining
        taken = np.zeros(n_nodes, bool)
        while cand.size:
            nmax = np.zeros(n_nodes, np.int64)
            np.maximum.at(nmax, cells[cand].ravel(), np.repeat(prio[cand], nloc))
            win = np.all(nmax[cells[cand]] == prio[cand][:, None], axis=1)
            chosen = cand[win]
            color[chosen] = col
            taken[cells[chosen].ravel()] = True
            cand = cand[~win]
            cand = cand[~np.any(taken[cells[cand]], axis=1)]
        remaining = remaining[color[remaining] < 0]
        col += 1
    return color, col


# ---------------------------------------------------------------------------
# face geometry / adjacency
# ---------------------------------------------------------------------------
def _face_local(dim):
    return ((0, 1), (1, 2), (2, 0)) if dim == 2 else ((0, 1, 2), (0, 1, 3), (1, 2, 3), (2, 0, 3))


def face_geometry(mesh: MeshArrays, faces: np.ndarray):
    """centre, unit normal and size of straight faces given by their corner nodes."""
    p = mesh.pos[faces[:, : mesh.dim]]
    centre = p.mean(axis=1)
    if mesh.dim == 2:
        t = p[:, 1] - p[:, 0]
        size = np.sqrt((t ** 2).sum(1))
        normal = np.stack([t[:, 1], -t[:, 0], np.zeros(len(t))], 1) / size[:, None]
    else:
        nvec = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
        nn = np.sqrt((nvec ** 2).sum(1))
        size = 0.5 * nn
        normal = nvec / nn[:, None]
    return centre, normal, size


def cell_adjacency(mesh: MeshArrays):
    """For every cell and local face: neighbour cell (-1 on the hull) and the face's corner nodes."""
    dim, N, C = mesh.dim, mesh.node_count, mesh.cell_count
    loc = _face_local(dim)
    nf = len(loc)
    c = mesh.cells[:, : mesh.nvert].astype(np.int64)
    fn = np.stack([c[:, list(l)] for l in loc], 1)            # (C, nf, dim)
    fs = np.sort(fn, axis=2)
    key = fs[..., 0]
    for j in range(1, dim):
        key = key * N + fs[..., j]
    key = key.ravel()
    order = np.argsort(key, kind="stable")
    ks = key[order]
    same_next = np.zeros(ks.size, bool)
    same_next[:-1] = ks[1:] == ks[:-1]
    nb = np.full(C * nf, -1, np.int64)
    i0 = order[:-1][same_next[:-1]]
    i1 = order[1:][same_next[:-1]]
    nb[i0] = i1 // nf
    nb[i1] = i0 // nf
    return nb.reshape(C, nf), fn


# ---------------------------------------------------------------------------
# reference element matrices on the unit simplex (exact rational values; the reference
# obtains the same numbers by quadrature that is exact for these degrees,
# elementmatrix.cpp:94-136, :706-781)
# ---------------------------------------------------------------------------
def unit_mass_matrix(kind: str) -> np.ndarray:
    """integral N_i N_j over an entity of unit size.  kind: edge2, edge3, tri3, tri6, tet4, tet10"""
    if kind == "edge2":
        return np.array([[2.0, 1.0], [1.0, 2.0]]) / 6.0
    if kind == "edge3":   # nodes: end, end, mid
        return np.array([[4.0, -1.0, 2.0], [-1.0, 4.0, 2.0], [2.0, 2.0, 16.0]]) / 30.0
    if kind == "tri3":
        return (np.ones((3, 3)) + np.eye(3)) / 12.0
    if kind == "tet4":
        return (np.ones((4, 4)) + np.eye(4)) / 20.0
    if kind == "tri6":    # corners 0,1,2; mids (0-1),(1-2),(2-0)
        M = np.zeros((6, 6))
        M[:3, :3] = -1.0
        M[np.arange(3), np.arange(3)] = 6.0
        M[3:, 3:] = 16.0
        M[np.arange(3, 6), np.arange(3, 6)] = 32.0
        opp = {0: 4, 1: 5, 2: 3}   # corner -> mid node of the opposite edge
        for v, mopp in opp.items():
            M[v, mopp] = M[mopp, v] = -4.0
        return M / 180.0
    if kind == "tet10":   # corners 0..3; mids (0-1),(0-2),(0-3),(1-2),(2-3),(3-1)
        edges = ((0, 1), (0, 2), (0, 3), (1, 2), (2, 3), (3, 1))
        M = np.zeros((10, 10))
        for i in range(4):
            for j in range(4):
                M[i, j] = 6.0 if i == j else 1.0
        for a, ea in enumerate(edges):
            for i in range(4):
                v = -4.0 if i in ea else -6.0
                M[i, 4 + a] = M[4 + a, i] = v
            for b, eb in enumerate(edges):
                if a == b:
                    M[4 + a, 4 + b] = 32.0
                elif set(ea) & set(eb):
                    M[4 + a, 4 + b] = 16.0
                else:
                    M[4 + a, 4 + b] = 8.0
        return M / 420.0
    raise KeyError(kind)


# ---------------------------------------------------------------------------
# mixed boundary condition coefficients
# ---------------------------------------------------------------------------
def mixed_bc_beta(centre, normal, source, k: float):
    """dcfemmodelling.cpp:430-506: mirror plane hard-wired at z (3-D) / y (2.5-D) = 0."""
    dimc = 1 if k > 0 else 2
    smir = np.array(source, float)
    smir[dimc] = -smir[dimc]
    r = source[None, :] - centre
    rm = smir[None, :] - centre
    ra = np.sqrt((r ** 2).sum(1))
    rma = np.sqrt((rm ** 2).sum(1))
    rn = np.abs((r * normal).sum(1))
    rmn = np.abs((rm * normal).sum(1))
    if k == 0:
        return ((rma * rma) * rn / ra + (ra * ra) * rmn / rma) / (rma * ra * (ra + rma))
    k0a, k0m = bessel_k0(ra * k), bessel_k0(rma * k)
    with np.errstate(divide="ignore", invalid="ignore"):
        res = k * (rn / ra * bessel_k1(ra * k) + rmn / rma * bessel_k1(rma * k)) / (k0a + k0m)
    return np.where((np.abs(k0a) < TOLERANCE) | (np.abs(k0m) < TOLERANCE), 0.0, res)


# ---------------------------------------------------------------------------
# shape functions (free electrodes)
# ---------------------------------------------------------------------------
def shape_functions(nloc: int, dim: int, L: np.ndarray) -> np.ndarray:
    """Lagrange shape functions at barycentric coordinates L (dim+1,) for P1/P2 simplices
    (core/tests/unittest/testFEM.h:91-140, :248-262 pin these forms)."""
    if nloc == dim + 1:
        return L.copy()
    edges = ((0, 1), (1, 2), (2, 0)) if dim == 2 else ((0, 1), (0, 2), (0, 3), (1, 2), (2, 3), (3, 1))
    return np.concatenate([L * (2.0 * L - 1.0), np.array([4.0 * L[a] * L[b] for a, b in edges])])


def locate_point(mesh: MeshArrays, p: np.ndarray):
    """containing cell + barycentric coordinates (brute force, set-up only)."""
    v = mesh.pos[mesh.cells[:, : mesh.nvert]][:, :, : mesh.dim]
    T = np.transpose(v[:, 1:] - v[:, :1], (0, 2, 1))
    rhs = (p[: mesh.dim][None, :] - v[:, 0])
    lam = np.linalg.solve(T, rhs[..., None])[..., 0]
    L = np.concatenate([1.0 - lam.sum(1, keepdims=True), lam], 1)
    ok = np.all(L >= -1e-10, axis=1)
    idx = np.nonzero(ok)[0]
    if idx.size == 0:
        return -1, None
    best = idx[np.argmax(L[idx].min(1))]
    return int(best), L[best]


# ---------------------------------------------------------------------------
# internal node renumbering (locality for the staged SpMM; invisible at the boundary)
# ---------------------------------------------------------------------------
def _spread_bits(v: np.ndarray, dim: int, bits: int) -> np.ndarray:
    out = np.zeros(v.shape, np.uint64)
    for b in range(bits):
        out |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(dim * b)
    return out


def node_ordering(mesh: MeshArrays) -> np.ndarray:
    """Space-filling-curve (Morton) order of the nodes: perm[new] = old.  Coordinates are
    rank-quantised per axis so graded meshes use the curve's resolution evenly."""
    dim = mesh.dim
    bits = 10 if dim == 3 else 15
    code = np.zeros(mesh.node_count, np.uint64)
    for ax in range(dim):
        u, inv = np.unique(mesh.pos[:, ax], return_inverse=True)
        if u.size <= (1 << bits):     # structured meshes: the rank itself, 2^d consecutive nodes = one grid cell (k_spmm_mma groups)
            q = inv.astype(np.uint64)
        else:
            q = (inv.astype(np.float64) * ((1 << bits) / max(1, u.size))).astype(np.uint64)
        code |= _spread_bits(q, dim, bits) << np.uint64(ax)
    return np.argsort(code, kind="stable").astype(np.int64)


def renumber_nodes(mesh: MeshArrays, perm: np.ndarray) -> MeshArrays:
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    return MeshArrays(mesh.dim, mesh.pos[perm], mesh.node_marker[perm], inv[mesh.cells].astype(np.int32),
                      mesh.cell_marker.copy(), inv[mesh.bounds].astype(np.int32) if mesh.bounds.size else mesh.bounds.copy(),
                      mesh.bound_marker.copy())


# ---------------------------------------------------------------------------
# the plan
# ---------------------------------------------------------------------------
class ERTPlan:
    """All geometry-derived arrays of one (mesh, scheme) pair; see build_plan.

    Node-indexed arrays are in the INTERNAL numbering (``node_perm[new] = old``); the reference's
    numbering is kept for everything that crosses the boundary: ``ref_rowptr/ref_colidx`` (the
    pattern exactly as SparseMatrix::buildSparsityPattern produces it), ``ref_slot`` (reference CSR
    slot -> internal slot) and ``node_inv`` (old -> new)."""


def build_plan(mesh: MeshArrays, scheme, k_values=None, weights=None, color_fn=None, reorder: bool = True) -> ERTPlan:
    P = ERTPlan()
    P.mesh_ref = mesh
    perm = node_ordering(mesh) if reorder else np.arange(mesh.node_count, dtype=np.int64)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    P.node_perm, P.node_inv = perm, inv
    src_order_ref = np.nonzero(mesh.node_marker == MARKER_NODE_ELECTRODE)[0]      # reference visiting order
    if reorder:
        mesh = renumber_nodes(mesh, perm)
    dim, N, C, nloc = mesh.dim, mesh.node_count, mesh.cell_count, mesh.nloc
    P.dim, P.N, P.C, P.nloc = dim, N, C, nloc
    P.mesh, P.scheme = mesh, scheme

    # ---- boundary classification / topography (dcfemmodelling.cpp:725-765) ---------
    bm = mesh.bound_marker
    neumann_domain = not np.any((bm == MARKER_BOUND_MIXED) | (bm == MARKER_BOUND_DIRICHLET))
    surf = np.nonzero(bm == MARKER_BOUND_NEUMANN)[0]
    topography = False
    surface_z = -MAX_DOUBLE
    if surf.size:
        # the reference compares Boundary::center()[dim-1] exactly (!=); centre = mean of all face nodes
        cz = mesh.pos[mesh.bounds[surf]].mean(axis=1)[:, dim - 1]
        surface_z = float(cz[0])
        topography = bool(np.any(cz != surface_z))
    if neumann_domain:
        topography = True
        if dim == 2:
            neumann_domain = False
    P.topography, P.surface_z, P.neumann_domain = topography, surface_z, neumann_domain
    # reference-electrode node (-999): the first one in the reference's node order (:1009-1015); a pure-Neumann domain
    # without one takes the last electrode as current reference (:1054-1064)
    ref_nodes = np.nonzero(P.mesh_ref.node_marker == MARKER_NODE_REFERENCE)[0]
    P.ref_node = int(inv[ref_nodes[0]]) if ref_nodes.size else -1
    P.ref_last = 1 if (neumann_domain and P.ref_node < 0) else 0
    # topography: the plan is the same; analytic primary potentials / analytic branches are switched off in the library
    # and CoreB200 supplies numeric primary potentials from a P2 total-field solve (dcfemmodelling.cpp:2009-2056)

    # ---- pattern, scatter map, colours -------------------------------------------
    P.rowptr, P.colidx, pos = build_pattern(mesh)
    P.nnz = int(P.colidx.size)
    if color_fn is None:
        color, ncol = color_cells_numpy(mesh.cells, N)
    else:
        color, ncol = color_fn(mesh.cells, N)
    order = np.argsort(color, kind="stable").astype(np.int32)      # colour-major cell order
    P.color_order = order
    P.color_ptr = np.concatenate([[0], np.cumsum(np.bincount(color, minlength=ncol))]).astype(np.int32)
    P.n_colors = int(ncol)
    P.cells_col = np.ascontiguousarray(mesh.cells[order].T)        # (nloc, C) SoA, colour order
    P.pos_col = np.ascontiguousarray(pos[order].T)                 # (nloc^2, C)
    P.diag_pos = csr_positions(P.rowptr, P.colidx, np.arange(N), np.arange(N))
    # the reference's pattern (original numbering) and the slot correspondence
    rowof_i = np.repeat(np.arange(N, dtype=np.int64), np.diff(P.rowptr))
    key_ref = perm[rowof_i] * N + perm[P.colidx]
    o = np.argsort(key_ref, kind="stable")
    P.ref_slot = o.astype(np.int64)
    P.ref_colidx = perm[P.colidx][o].astype(np.int32)
    P.ref_rowptr = np.concatenate([[0], np.cumsum(np.bincount(perm[rowof_i], minlength=N))]).astype(np.int32)

    # ---- electrodes (dcfemmodelling.cpp:845-940) ------------------------------------
    sens = np.array(scheme.sensors, float).reshape(-1, 3)
    if dim == 2:
        zvar = np.ptp(sens[:, 2]) > 0 or np.max(np.abs(sens[:, 2])) > 0
        yflat = (np.ptp(sens[:, 1]) == 0) and np.max(np.abs(sens[:, 1])) < 1e-8
        if zvar and yflat:
            sens = sens[:, [0, 2, 1]].copy()
    nE = sens.shape[0]
    P.nE = nE
    src_nodes = [int(i) for i in inv[src_order_ref]]          # candidates in the reference's node order
    el_node = np.full(nE, -1, np.int64)          # mID: node id for node electrodes
    el_pos = np.zeros((nE, 3))
    el_cell = np.full(nE, -1, np.int64)
    pick_ptr, pick_idx, pick_w = [0], [], []      # potential pick-up / RHS stencil
    sing_node = np.full(nE, -1, np.int64)        # node whose analytic value gets patched
    for i in range(nE):
        hit = -1
        for t, nidx in enumerate(src_nodes):
            if np.sqrt(((sens[i] - mesh.pos[nidx]) ** 2).sum()) < 0.01:
                hit = t
                break
        if hit >= 0:
            nidx = src_nodes.pop(hit)
            el_node[i] = nidx
            el_pos[i] = mesh.pos[nidx]
            sing_node[i] = nidx
            pick_idx.append([nidx])
            pick_w.append([1.0])
        else:
            cidx, L = locate_point(mesh, sens[i])
            if cidx < 0:
                raise ValueError("There is a requested electrode that does not match the given mesh.")
            el_cell[i] = cidx
            el_pos[i] = sens[i]
            ids = mesh.cells[cidx]
            pick_idx.append(list(ids))
            pick_w.append(list(shape_functions(nloc, dim, L)))
            d = np.sqrt(((mesh.pos[ids] - sens[i]) ** 2).sum(1))
            near = np.nonzero(d < 1e-4)[0]
            if near.size:
                sing_node[i] = ids[near[-1]]
        pick_ptr.append(pick_ptr[-1] + len(pick_idx[-1]))
    P.el_node, P.el_pos, P.el_cell, P.sing_node = el_node, el_pos, el_cell, sing_node
    P.el_node_ref = np.where(el_node >= 0, perm[np.maximum(el_node, 0)], -1)
    P.pick_ptr = np.asarray(pick_ptr, np.int32)
    P.pick_idx = np.concatenate(pick_idx).astype(np.int32)
    P.pick_w = np.concatenate(pick_w).astype(np.float64)
    P.source_center = el_pos.sum(0) / float(nE)

    # node -> cells incidence for the electrodes (rho at the source, electrode.cpp:102-120, :252-268)
    flat = mesh.cells.ravel()
    order_nc = np.argsort(flat, kind="stable")
    nc_ptr = np.concatenate([[0], np.cumsum(np.bincount(flat, minlength=N))])
    nc_cells = (order_nc // nloc)

    def cells_of(node):
        return np.unique(nc_cells[nc_ptr[node]:nc_ptr[node + 1]])

    inc_ptr, inc_cells = [0], []
    min_radius = np.zeros(nE)
    for i in range(nE):
        cs = cells_of(el_node[i]) if el_node[i] >= 0 else np.array([el_cell[i]])
        inc_cells.append(cs)
        inc_ptr.append(inc_ptr[-1] + cs.size)
        if sing_node[i] >= 0:
            cs2 = cells_of(sing_node[i])
            nbn = np.unique(mesh.cells[cs2].ravel())
            nbn = nbn[nbn != sing_node[i]]
            min_radius[i] = np.sqrt(((mesh.pos[nbn] - mesh.pos[sing_node[i]]) ** 2).sum(1)).min()
    P.src_cell_ptr = np.asarray(inc_ptr, np.int32)
    P.src_cells = np.concatenate(inc_cells).astype(np.int32)
    P.min_radius = min_radius

    # ---- wavenumbers ---------------------------------------------------------------
    if k_values is not None and weights is not None:
        P.k, P.w = np.asarray(k_values, float).copy(), np.asarray(weights, float).copy()
    else:
        P.k, P.w = init_kwave_list(dim, scheme.sensors)
    P.nK = int(P.k.size)
    P.nS = P.nK * nE

    # singular-value patch per (electrode, k) (electrode.cpp:154-189 with scale = 0)
    sing_val = np.zeros((P.nK, nE))
    for kk, kv in enumerate(P.k):
        if kv > 0.0:
            with np.errstate(divide="ignore", invalid="ignore"):
                sing_val[kk] = bessel_k0(min_radius / 6.0 * kv) / math.pi
        else:
            with np.errstate(divide="ignore"):
                sing_val[kk] = 1.0 / (2.0 * math.pi * min_radius / 2.0)
    sing_val[:, sing_node < 0] = 0.0
    P.sing_val = sing_val

    # ---- mixed boundary faces -> (csr slot, owner cell, coef[k]) triplets -----------
    owner = mesh.bound_owner() if mesh.bounds.shape[0] else np.zeros(0, np.int32)
    mixed = np.nonzero(bm == MARKER_BOUND_MIXED)[0]
    nlb = mesh.bounds.shape[1] if mesh.bounds.ndim == 2 else 0
    if mixed.size:
        fb = mesh.bounds[mixed]
        centre, normal, size = face_geometry(mesh, fb)
        kind = {(2, 2): "edge2", (2, 3): "edge3", (3, 3): "tri3", (3, 6): "tri6"}[(dim, nlb)]
        U = unit_mass_matrix(kind)
        rows = np.repeat(fb, nlb, axis=1).ravel()
        cols = np.tile(fb, (1, nlb)).ravel()
        slot = csr_positions(P.rowptr, P.colidx, rows, cols)
        own = np.repeat(owner[mixed], nlb * nlb)
        coef = np.zeros((P.nK, slot.size))
        for kk, kv in enumerate(P.k):
            beta = mixed_bc_beta(centre, normal, P.source_center, float(kv))
            coef[kk] = (beta * size)[:, None].repeat(nlb * nlb, 1).ravel() * np.tile(U.ravel(), fb.shape[0])
        o = np.argsort(slot, kind="stable")
        slot, own, coef = slot[o], own[o], coef[:, o]
        uslot, start = np.unique(slot, return_index=True)
        P.bc_slot = uslot.astype(np.int32)
        P.bc_ptr = np.concatenate([start, [slot.size]]).astype(np.int32)
        P.bc_owner = own.astype(np.int32)
        P.bc_coef = np.ascontiguousarray(coef)
    else:
        P.bc_slot = np.zeros(0, np.int32)
        P.bc_ptr = np.zeros(1, np.int32)
        P.bc_owner = np.zeros(0, np.int32)
        P.bc_coef = np.zeros((P.nK, 0))

    # ---- homogeneous Dirichlet nodes (-3 faces; calibration nodes are ignored on
    #      non-Neumann domains, dcfemmodelling.cpp:1066-1070) -------------------------
    dn = np.unique(mesh.bounds[bm == MARKER_BOUND_DIRICHLET].ravel()) if np.any(bm == MARKER_BOUND_DIRICHLET) else np.zeros(0, np.int64)
    if neumann_domain:
        # calibration nodes (-1000) pin the potential of a pure-Neumann domain; without one the reference takes its node 0
        # (dcfemmodelling.cpp:1044-1052)
        cal = np.nonzero(mesh.node_marker == MARKER_NODE_CALIBRATION)[0]
        if cal.size == 0:
            cal = np.array([inv[0]], np.int64)
        dn = np.unique(np.concatenate([dn, cal]))
    P.dir_nodes = dn.astype(np.int32)
    rowof = np.repeat(np.arange(N, dtype=np.int64), np.diff(P.rowptr))
    isd = np.zeros(N, bool)
    isd[dn] = True
    kill = isd[rowof] | isd[P.colidx]
    P.dir_zero_slots = np.nonzero(kill)[0].astype(np.int32)
    P.dir_diag_slots = P.diag_pos[dn].astype(np.int32) if dn.size else np.zeros(0, np.int32)

    # ---- model mapping (modellingbase.cpp:401-497, mesh.cpp:2247-2316) ---------------
    cm = mesh.cell_marker
    if np.any(cm <= -1000000):
        raise NotImplementedError("fixed-value regions are not supported on the B200 path")
    P.M = int(cm.max()) + 1 if cm.size and cm.max() >= 0 else 0
    P.cell_marker = cm.copy()
    # every cell without a model value is filled by the prolongation: the reference zero-initialises the attributes and
    # prolongateEmptyCellsValues (mesh.cpp:2247-2316) fills all |value| < TOLERANCE, whatever negative marker they carry
    empty = cm < 0
    P.has_background = bool(empty.any())
    levels = []
    if P.has_background:
        nb, fn = cell_adjacency(mesh)
        nf = nb.shape[1]
        _, fnormal, _ = face_geometry(mesh, fn.reshape(-1, dim).astype(np.int64))
        xy = np.array([1.0, 1.0, 0.0]) if dim == 3 else np.array([1.0, 0.0, 0.0])
        zw = (np.sqrt(((fnormal * xy) ** 2).sum(1)) + 1e-6).reshape(C, nf)
        filled = ~empty
        lvl = np.where(filled, 0, -1)
        cur = 0
        while True:
            todo = np.nonzero(lvl < 0)[0]
            if todo.size == 0:
                break
            nbt = nb[todo]
            ok = (nbt >= 0) & (lvl[np.maximum(nbt, 0)] >= 0) & (lvl[np.maximum(nbt, 0)] <= cur)
            has = ok.any(1)
            if not has.any():
                raise RuntimeError("cannot fill empty cells: disconnected background region")
            cur += 1
            cells_l = todo[has]
            wts = np.where(ok[has], zw[cells_l], 0.0)
            wts = wts / wts.sum(1, keepdims=True)
            levels.append((cells_l.astype(np.int32), np.where(ok[has], nbt[has], 0).astype(np.int32), wts))
            lvl[cells_l] = cur
        P.pro_nf = nf
    P.pro_levels = levels

    # ---- Jacobian columns: cells with marker >= 0 sorted by marker (bertJacobian.cpp:298-299)
    para = np.nonzero(cm >= 0)[0]
    o = np.argsort(cm[para], kind="stable")
    P.jac_cells = para[o].astype(np.int32)
    P.jac_col_ptr = np.concatenate([[0], np.cumsum(np.bincount(cm[para], minlength=P.M))]).astype(np.int32)
    return P
