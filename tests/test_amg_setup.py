"""Host-side set-up of the multilevel preconditioner (pygimli_b200/amg_setup.py): structural invariants of the
aggregation hierarchy and a numpy emulation of the GPU V-cycle showing that it is a symmetric positive definite
preconditioner that cuts the PCG iteration count.  CPU-only."""
import numpy as np
import pytest
import scipy.sparse as sp

from cases import make_case
from pygimli_b200 import _capi, amg_setup, host_setup as hs


def _matrix(name):
    """rho = 1 stiffness matrix of a test mesh assembled with numpy (P1 closed form), internal numbering"""
    mesh, scheme, _ = make_case(name)
    P = hs.build_plan(mesh, scheme, color_fn=_capi.color_cells)
    m = P.mesh
    X = m.pos[m.cells[:, : m.dim + 1], : m.dim]
    J = np.transpose(X[:, 1:] - X[:, :1], (0, 2, 1))
    size = np.abs(np.linalg.det(J)) / (2.0 if m.dim == 2 else 6.0)
    Jinv = np.linalg.inv(J)                                          # rows: grad of lambda_1..d
    g = np.concatenate([-Jinv.sum(1, keepdims=True), Jinv], 1)       # (C, d+1, d)
    K = size[:, None, None] * np.einsum("cid,cjd->cij", g, g)
    rows = np.repeat(m.cells, m.nloc, axis=1).ravel()
    cols = np.tile(m.cells, (1, m.nloc)).ravel()
    A = sp.coo_matrix((K.ravel(), (rows, cols)), shape=(P.N, P.N)).tocsr()
    A = A + sp.diags(np.full(P.N, 1e-3 * A.diagonal().mean()))       # mixed-BC stand-in: makes it definite
    A.sort_indices()
    return P, A


@pytest.fixture(scope="module", params=["2d_p1", "3d_p1"])
def hier(request):
    P, A = _matrix(request.param)
    lv = amg_setup.build_hierarchy(A.indptr, A.indices, A.data, _capi.pairwise_aggregate, min_size=64)
    return P, A, lv


def test_hierarchy_structure(hier):
    P, A, lv = hier
    assert len(lv) >= 2
    n_f, nnz_f = P.N, A.nnz
    for L in lv:
        assert L["n"] < 0.7 * n_f                                    # real coarsening
        assert L["agg"].size == n_f and L["agg"].min() == 0 and L["agg"].max() == L["n"] - 1
        assert np.array_equal(np.sort(L["mem_idx"]), np.arange(n_f))            # every fine node in exactly one aggregate
        assert np.array_equal(L["agg"][L["mem_idx"]], np.repeat(np.arange(L["n"]), np.diff(L["mem_ptr"])))
        assert L["gal_ptr"].size == L["nnz"] + 1 and L["gal_idx"].size == nnz_f
        assert np.array_equal(np.sort(L["gal_idx"]), np.arange(nnz_f))          # every fine entry summed exactly once
        rowof = np.repeat(np.arange(L["n"]), np.diff(L["rowptr"]))
        assert np.array_equal(L["colidx"][L["diag_pos"]], np.arange(L["n"])) and np.array_equal(rowof[L["diag_pos"]], np.arange(L["n"]))
        n_f, nnz_f = L["n"], L["nnz"]


def test_galerkin_gather_equals_triple_product(hier):
    P, A, lv = hier
    L = lv[0]
    vc = amg_setup._sum_values(A.data, L["gal_ptr"], L["gal_idx"])
    Ac = sp.csr_matrix((vc, L["colidx"], L["rowptr"]), shape=(L["n"], L["n"]))
    Pm = sp.csr_matrix((np.ones(P.N), (np.arange(P.N), L["agg"])), shape=(P.N, L["n"]))
    ref = (Pm.T @ A @ Pm).tocsr()
    assert abs(Ac - ref).max() < 1e-12 * abs(ref).max()


def _vcycle(mats, lv, l, r, sweeps=8):
    A = mats[l]
    g = (abs(A).sum(1).A1 / A.diagonal()).max()
    dw = (1.6 / max(2.0, g)) / A.diagonal()
    if l == len(lv):
        x = dw * r
        for _ in range(sweeps - 1):
            x = x + dw * (r - A @ x)
        return x
    L = lv[l]
    x = dw * r
    rc = np.add.reduceat((r - A @ x)[L["mem_idx"]], L["mem_ptr"][:-1])
    x = x + _vcycle(mats, lv, l + 1, rc, sweeps)[L["agg"]]
    return x + dw * (r - A @ x)


def test_vcycle_is_spd_and_accelerates_pcg(hier):
    P, A, lv = hier
    mats, v = [A], A.data
    for L in lv:
        v = amg_setup._sum_values(v, L["gal_ptr"], L["gal_idx"])
        mats.append(sp.csr_matrix((v, L["colidx"], L["rowptr"]), shape=(L["n"], L["n"])))
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(P.N), rng.standard_normal(P.N)
    Ma, Mb = _vcycle(mats, lv, 0, a), _vcycle(mats, lv, 0, b)
    assert abs(a @ Mb - b @ Ma) < 1e-10 * abs(a @ Mb)              # symmetric
    assert a @ Ma > 0 and b @ Mb > 0                                 # positive

    def pcg(Minv):
        x = np.zeros(P.N); r = b.copy(); z = Minv(r); p = z.copy(); rz = r @ z
        for it in range(1, 5000):
            Ap = A @ p; al = rz / (p @ Ap); x += al * p; r -= al * Ap
            if np.sqrt(r @ r) <= 1e-10 * np.sqrt(b @ b):
                return it
            z = Minv(r); rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn
        return 5000
    it_jac = pcg(lambda r: r / A.diagonal())
    it_ml = pcg(lambda r: _vcycle(mats, lv, 0, r))
    assert it_ml < 0.5 * it_jac


def _aniso_grid(nx, ny, sx, sy):
    """5-point matrix of -sx u_xx - sy u_yy on an nx x ny grid (+ a small shift): couplings -sx along x, -sy along y"""
    idx = np.arange(nx * ny).reshape(ny, nx)
    rows, cols, vals = [], [], []
    for (a, b, s) in ((idx[:, :-1], idx[:, 1:], sx), (idx[:-1, :], idx[1:, :], sy)):
        a, b = a.ravel(), b.ravel()
        rows += [a, b]; cols += [b, a]; vals += [np.full(a.size, -s), np.full(a.size, -s)]
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nx * ny, nx * ny)).tocsr()
    A = A + sp.diags(-A.sum(1).A1 + 1e-3)
    A.sort_indices()
    return A, idx


def test_strength_threshold_keeps_aggregates_in_the_strong_direction():
    """stretched cells: with the threshold no aggregate crosses the weak (y) couplings, however many passes follow;
    the unconditional matching (theta = 0) pairs across them as soon as the strong partners are used up"""
    A, idx = _aniso_grid(9, 8, 100.0, 1.0)                           # odd row length: one node per row is left over
    row_of = np.repeat(np.arange(8), 9)
    agg, na = _capi.pairwise_aggregate(A.indptr, A.indices, A.data)               # default theta = 0.25
    for a in range(na):
        assert np.unique(row_of[agg == a]).size == 1                 # members share a grid row
    assert np.bincount(agg).max() <= 3                               # pairs, the left-over node joins one of them
    agg0, na0 = _capi.pairwise_aggregate(A.indptr, A.indices, A.data, theta=0.0)
    assert np.bincount(agg0).min() >= 2                              # nobody stays alone without a threshold


def test_weak_partner_is_not_matched_but_may_join():
    """a coupling must be strong for BOTH nodes to form a pair; a left-over node joins across a coupling that is strong
    for itself"""
    A = sp.lil_matrix((5, 5))
    for i, j, s in ((0, 1, 10.0), (2, 3, 10.0), (1, 4, 0.1), (3, 4, 0.1)):      # node 4 hangs on two strong pairs
        A[i, j] = A[j, i] = -s
    A.setdiag(-np.asarray(A.sum(1)).ravel() + 1e-3)
    A = A.tocsr(); A.sort_indices()
    agg, na = _capi.pairwise_aggregate(A.indptr, A.indices, A.data)
    assert na == 2
    # node 4's couplings (0.1) are strong for ITSELF (they are all it has) -> it joins a neighbouring aggregate
    assert np.sum(agg == agg[4]) == 3
    # but it is never chosen as a PARTNER: 0.1 < 0.25 * 10 for nodes 1 and 3
    assert agg[0] == agg[1] and agg[2] == agg[3]


def test_threshold_improves_the_preconditioner_on_a_graded_mesh():
    """numpy emulation of the V-cycle: the thresholded matching needs fewer PCG iterations than the unconditional one
    on the (graded, stretched-cell) 3-D test mesh"""
    P, A = _matrix("3d_p1")
    counts = {}
    for theta in (0.0, 0.25):
        lv = amg_setup.build_hierarchy(A.indptr, A.indices, A.data,
                                       lambda rp, ci, v, theta=theta: _capi.pairwise_aggregate(rp, ci, v, theta=theta), min_size=64)
        mats, v = [A], A.data
        for L in lv:
            v = amg_setup._sum_values(v, L["gal_ptr"], L["gal_idx"])
            mats.append(sp.csr_matrix((v, L["colidx"], L["rowptr"]), shape=(L["n"], L["n"])))
        dws = [(1.6 / max(2.0, (abs(M).sum(1).A1 / M.diagonal()).max())) / M.diagonal() for M in mats]

        def vc(l, r):
            Al, dw = mats[l], dws[l]
            if l == len(lv):
                x = dw * r
                for _ in range(7):
                    x = x + dw * (r - Al @ x)
                return x
            L = lv[l]
            x = dw * r
            x = x + vc(l + 1, np.add.reduceat((r - Al @ x)[L["mem_idx"]], L["mem_ptr"][:-1]))[L["agg"]]
            return x + dw * (r - Al @ x)
        b = np.random.default_rng(0).standard_normal(P.N)
        x, r = np.zeros(P.N), b.copy()
        z = vc(0, r); p = z.copy(); rz = r @ z
        it = 0
        while np.linalg.norm(r) > 1e-10 * np.linalg.norm(b) and it < 2000:
            Ap = A @ p; al = rz / (p @ Ap); x += al * p; r -= al * Ap
            z = vc(0, r); rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn; it += 1
        counts[theta] = it
    assert counts[0.25] <= counts[0.0]
    assert counts[0.25] < 2000
