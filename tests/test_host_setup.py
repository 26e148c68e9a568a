"""Host-side plan building (geometry-only set-up of the product) against the reference's golden
vectors and against the oracle restatement.  CPU-only."""
import os

import numpy as np
import pytest

from cases import CASES, make_case
from pygimli_b200 import _capi, host_setup as hs
from pygimli_b200.mesh import create_p2, create_h2, grid_mesh_2d, grid_mesh_3d
from pygimli_b200.scheme import create_dd, create_slm, create_dd_complete, geometric_factors

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module", params=CASES)
def plan(request):
    mesh, scheme, model = make_case(request.param)
    P = hs.build_plan(mesh, scheme, color_fn=_capi.color_cells)
    return request.param, mesh, scheme, model, P, np.load(os.path.join(GOLD, request.param + ".npz"))


def test_reference_pattern_bit_exact(plan):
    _, _, _, _, P, g = plan
    assert np.array_equal(P.ref_rowptr, g["rowptr"]) and np.array_equal(P.ref_colidx, g["colidx"])
    assert P.ref_rowptr.dtype == np.int32 and P.ref_colidx.dtype == np.int32


def test_wavenumbers_and_electrodes(plan):
    _, _, _, _, P, g = plan
    assert np.array_equal(P.k, g["k"]) and np.array_equal(P.w, g["w"])
    assert np.array_equal(P.el_node_ref, g["el_nodes"])


def test_internal_numbering_is_a_permutation(plan):
    _, mesh, _, _, P, _ = plan
    assert np.array_equal(np.sort(P.node_perm), np.arange(mesh.node_count))
    assert np.allclose(P.mesh.pos, mesh.pos[P.node_perm])
    # the scatter map points at the right (row, col)
    rowof = np.repeat(np.arange(P.N), np.diff(P.rowptr))
    nl = P.nloc
    cells_col = P.cells_col.T                       # (C, nloc) colour order
    pos = P.pos_col.T                               # (C, nloc^2)
    rows = np.repeat(cells_col, nl, axis=1)
    cols = np.tile(cells_col, (1, nl))
    assert np.array_equal(rowof[pos], rows) and np.array_equal(P.colidx[pos], cols)


def test_colours_are_conflict_free(plan):
    _, _, _, _, P, _ = plan
    cells_col = P.cells_col.T
    for c in range(P.n_colors):
        nodes = cells_col[P.color_ptr[c]:P.color_ptr[c + 1]].ravel()
        assert np.unique(nodes).size == nodes.size


def test_prolongation_levels_reproduce_reference_mapping(plan):
    """emulate the per-level GPU kernel in numpy: the mapped model must equal the reference's"""
    name, mesh, _, model, P, g = plan
    if model.size == mesh.cell_count:
        rho = model.copy()
    else:
        rho = np.where(mesh.cell_marker >= 0, model[np.maximum(mesh.cell_marker, 0)], 0.0)
        for cells, nb, w in P.pro_levels:
            rho[cells] = (w * rho[nb]).sum(1)
    assert np.max(np.abs(rho - g["rho"]) / g["rho"]) < 1e-13


def test_jacobian_columns(plan):
    _, mesh, _, _, P, _ = plan
    cm = mesh.cell_marker
    assert P.M == cm.max() + 1
    assert np.all(np.diff(cm[P.jac_cells]) >= 0) and np.all(cm[P.jac_cells] >= 0)
    assert P.jac_col_ptr[-1] == (cm >= 0).sum()


def test_scheme_sizes_match_reference_generators():
    s41 = np.zeros((41, 3)); s41[:, 0] = np.arange(41)
    s96 = np.zeros((96, 3)); s96[:, 0] = np.arange(96)
    assert create_dd(s41).size == 741
    assert create_slm(s96).size == 2209
    s100 = np.zeros((100, 3)); s100[:, 0] = np.arange(100)
    assert create_dd_complete(s100).size == 9700


def test_geometric_factor_known_value():
    # Wenner alpha with spacing a on a flat half-space: k = 2 pi a
    s = np.zeros((4, 3)); s[:, 0] = np.arange(4) * 2.0
    from pygimli_b200.scheme import SchemeArrays
    sch = SchemeArrays(s, [0], [3], [1], [2])
    assert abs(geometric_factors(sch, 3)[0] - 2 * np.pi * 2.0) < 1e-12


def test_refinement_against_reference_if_built():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    m2 = grid_mesh_2d(np.linspace(0, 3, 4), -np.linspace(0, 2, 3))
    m3 = grid_mesh_3d(np.linspace(0, 2, 3), np.linspace(0, 2, 3), -np.linspace(0, 2, 3))
    for m in (m2, m3):
        for kind, mine in ((2, create_p2(m)), (1, create_h2(m))):
            r = ref.refine(m, kind)
            assert r["cells"].shape == mine.cells.shape
            # same cell-local geometry: node positions per cell-local index agree (P2), same volumes (H2)
            if kind == 2:
                assert np.allclose(r["pos"][r["cells"]], mine.pos[mine.cells])
            else:
                assert np.isclose(mine.cell_sizes().sum(), m.cell_sizes().sum())
                assert sorted(np.round(mine.cell_sizes(), 12)) == pytest.approx(sorted(np.round(
                    __import__("pygimli_b200").MeshArrays(m.dim, r["pos"], r["node_marker"], r["cells"], r["cell_marker"],
                                                          r["bounds"], r["bound_marker"]).cell_sizes(), 12)))
            assert np.array_equal(np.sort(r["cell_marker"]), np.sort(mine.cell_marker))
