"""The compiled plan builder (csrc/plan_builder.cpp: pgb200_plan_build / pgb200_plan_build_hierarchy, what a reference-side
binding calls with a mesh and a data container) against its numpy twin (pygimli_b200/host_setup.py, amg_setup.py), which
the golden-vector tests pin to the reference: every index array bit-exact, floating-point tables to 1e-14 relative
(libm vs numpy transcendental functions), wavenumbers bit-exact.  CPU only."""
import numpy as np
import pytest

from cases import CASES, WIDE_CASES, TOPO_CASES, make_case, make_wide_case, make_topo_case, make_pole_case, make_tank_case
from pygimli_b200 import _capi, amg_setup, host_setup as hs
from pygimli_b200.scheme import geometric_factors

INT_ARRAYS = ["node_perm", "node_inv", "rowptr", "colidx", "diag_pos", "ref_rowptr", "ref_colidx", "ref_slot", "color_ptr",
              "color_order", "bc_slot", "bc_ptr", "bc_owner", "dir_zero_slots", "dir_diag_slots", "dir_nodes", "sing_node",
              "pick_ptr", "pick_idx", "src_cell_ptr", "src_cells", "jac_cells", "jac_col_ptr", "el_node", "el_cell"]


def _variants():
    for name in CASES:
        yield name, make_case(name)[:2], {}
    for name in WIDE_CASES[:2]:
        yield name, make_wide_case(name)[:2], {}
    for name in TOPO_CASES:
        yield name, make_topo_case(name)[:2], {}
    yield "pole", make_pole_case()[:2], {}
    mesh, scheme, _ = make_case("3d_p1")                      # -3 faces: homogeneous Dirichlet rows
    zb = mesh.pos[mesh.bounds].mean(1)[:, 2]
    mesh.bound_marker[np.isclose(zb, mesh.pos[:, 2].min())] = -3
    yield "dirichlet", (mesh, scheme), {}
    mesh, scheme, _ = make_case("2d_p2")                      # free electrodes (no electrode nodes)
    mesh.node_marker[:] = 0
    sch = scheme.subset(np.arange(scheme.size))
    sch.sensors = scheme.sensors.copy()
    sch.sensors[:, 0] += 0.13
    sch.k = geometric_factors(sch, 2)
    yield "free", (mesh, sch), {}
    for nm, flags in (("tank_ref_cal", (True, True)), ("tank_ref", (True, False)), ("tank_last", (False, True))):
        mesh, scheme, _ = make_tank_case(*flags)              # pure-Neumann tank: calibration node, reference electrode
        yield nm, (mesh, scheme), {}
    mesh, scheme, _ = make_case("2d_p1")                      # user wavenumbers
    yield "user_k", (mesh, scheme), dict(k_values=np.array([0.01, 0.05, 0.2, 0.8, 2.5]), weights=np.array([0.02, 0.05, 0.2, 0.6, 1.1]))


@pytest.mark.parametrize("name,ms,kw", list(_variants()), ids=[v[0] for v in _variants()])
def test_native_plan_equals_numpy_twin(name, ms, kw):
    mesh, scheme = ms
    P = hs.build_plan(mesh, scheme, kw.get("k_values"), kw.get("weights"), color_fn=_capi.color_cells)
    Q = _capi.plan_build(mesh, scheme, True, kw.get("k_values"), kw.get("weights"))
    for s in ("N", "C", "nnz", "nE", "nK", "M", "dim", "nloc", "n_colors"):
        assert getattr(Q, s) == getattr(P, s), s
    assert Q.topography == P.topography and Q.has_background == P.has_background
    assert Q.ref_node == P.ref_node and Q.ref_last == P.ref_last and Q.neumann_domain == P.neumann_domain
    for nm in INT_ARRAYS:
        a, b = np.asarray(getattr(P, nm)), Q.array(nm)
        assert a.shape == b.shape and np.array_equal(a.astype(np.int64), b.astype(np.int64)), nm
    assert np.array_equal(P.cells_col.ravel(), Q.array("cells_col")) and np.array_equal(P.pos_col.ravel(), Q.array("pos_col"))
    assert np.array_equal(P.k, Q.array("k")) and np.array_equal(P.w, Q.array("w"))          # bit-exact, as against the reference
    for nm, ref in (("bc_coef", P.bc_coef.ravel()), ("sing_val", P.sing_val.ravel()), ("pick_w", P.pick_w),
                    ("el_pos", P.el_pos.ravel()), ("min_radius", P.min_radius)):
        b = Q.array(nm)
        assert b.shape == ref.shape, nm
        if ref.size:
            assert np.max(np.abs(b - ref)) <= 1e-14 * max(np.max(np.abs(ref)), 1e-300), nm
    lv = P.pro_levels
    if lv:
        assert np.array_equal(np.concatenate([c for c, _, _ in lv]), Q.array("pro_cells"))
        assert np.array_equal(np.concatenate([n for _, n, _ in lv]).ravel(), Q.array("pro_nb"))
        assert np.max(np.abs(np.concatenate([w for _, _, w in lv]).ravel() - Q.array("pro_w"))) <= 1e-15
        assert np.array_equal(np.cumsum([0] + [len(c) for c, _, _ in lv]), Q.array("pro_level_ptr"))
    if scheme.k is None and not P.topography:
        assert np.max(np.abs(Q.array("k_fac") - geometric_factors(scheme, mesh.dim)) / np.abs(geometric_factors(scheme, mesh.dim))) < 1e-14
    Q.free()


@pytest.mark.parametrize("name", ["3d_p1", "2d_p2", "3d_p1_wide"])
def test_native_hierarchy_equals_numpy_twin(name):
    mesh, scheme = (make_wide_case(name) if name.endswith("wide") else make_case(name))[:2]
    Q = _capi.plan_build(mesh, scheme)
    rp, ci, N = Q.rowptr, Q.colidx, Q.N
    pos = Q.array("pos").reshape(-1, 3)
    rowof = np.repeat(np.arange(N), np.diff(rp))
    dist = np.linalg.norm(pos[rowof] - pos[ci], axis=1)
    vals = np.where(rowof == ci, 0.0, -1.0 / np.maximum(dist, 1e-9) ** 2)          # an M-matrix with graded couplings
    vals[rowof == ci] = np.bincount(rowof, weights=-vals, minlength=N) * 1.001
    vals = np.ascontiguousarray(vals)
    nl = _capi.lib().pgb200_plan_build_hierarchy(Q._ptr, vals.ctypes.data, 0.25, 2, 256, 12)
    lv = amg_setup.build_hierarchy(rp, ci, vals, _capi.pairwise_aggregate)
    assert nl == len(lv) and nl >= 1
    for l, L in enumerate(lv):
        for f in ("rowptr", "colidx", "diag_pos", "gal_ptr", "gal_idx", "agg", "mem_ptr", "mem_idx"):
            a, b = np.asarray(L[f]), Q.array(f"level{l}.{f}")
            assert a.shape == b.shape and np.array_equal(a, b), (l, f)
    Q.free()


def test_plan_builder_error_behaviour():
    mesh, scheme, _ = make_case("2d_p1")
    mesh.node_marker[:] = 0
    scheme.sensors[0, 0] = -1e6                    # an electrode outside the mesh
    with pytest.raises(ValueError, match="does not match the given mesh"):
        _capi.plan_build(mesh, scheme)
