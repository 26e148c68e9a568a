"""GPU parity tests: the CUDA path (through the C ABI) against the reference.

Checker = oracle/_ref (the reference C++ compiled from source; travels to the GPU box as a
prebuilt .so) when present, and always the committed golden vectors in tests/golden/ that
the reference produced in the build container (tests/make_golden.py).

Tolerances (BASELINE.json north_star): sparsity pattern and indexing bit-exact; matrix values
1e-12 relative to the row scale (summation order differs); potentials, apparent resistivities
and Jacobian 1e-8 relative (norm-wise per vector / matrix), rhoa additionally one quantum of
the reference's round(u, 1e-10): |k| * 1e-10.
"""
import os

import numpy as np
import pytest

from cases import CASES, make_case

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-8


def _fop(mesh, scheme, sr=True):
    from pygimli_b200 import ERTModellingB200
    fop = ERTModellingB200(sr=sr)
    fop.setMesh(mesh)
    fop.setData(scheme)
    return fop


def _relmax(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module", params=CASES)
def case(request):
    name = request.param
    mesh, scheme, model = make_case(name)
    g = np.load(os.path.join(GOLD, name + ".npz"))
    fop = _fop(mesh, scheme)
    rhoa = fop.response(model)
    return dict(name=name, mesh=mesh, scheme=scheme, model=model, g=g, fop=fop, rhoa=rhoa)


def test_pattern_bit_exact(case):
    P = case["fop"]._core._plan
    assert np.array_equal(P.ref_rowptr, case["g"]["rowptr"])
    assert np.array_equal(P.ref_colidx, case["g"]["colidx"])


def test_wavenumbers(case):
    P = case["fop"]._core._plan
    assert np.array_equal(P.k, case["g"]["k"])
    assert np.array_equal(P.w, case["g"]["w"])


def test_model_mapping(case):
    rho = case["fop"]._core.get("rho")
    assert _relmax(rho, case["g"]["rho"]) < 1e-13
    assert _relmax(case["fop"].mapERTModel(case["model"]), case["g"]["rho"]) < 1e-13


def test_matrix_values(case):
    core = case["fop"]._core
    P = core._plan
    vals = core.get("vals").reshape(P.nK, P.nnz)
    for kk, key in ((0, "vals_k0"), (P.nK - 1, "vals_klast")):
        ref = case["g"][key]
        rowof = np.repeat(np.arange(P.N), np.diff(P.ref_rowptr))
        scale = np.maximum.reduceat(np.abs(ref), P.ref_rowptr[:-1])[rowof]
        assert np.max(np.abs(vals[kk] - ref) / scale) < 1e-12


def test_potentials(case):
    core = case["fop"]._core
    P = core._plan
    pots = core.get("pots").reshape(P.nS, P.N)
    for r, ref in zip(case["g"]["pots_rows"], case["g"]["pots"]):
        assert _relmax(pots[r], ref) < TOL


def test_apparent_resistivity(case):
    ref = case["g"]["rhoa"]
    kf = np.abs(case["g"]["kfac"])
    err = np.abs(case["rhoa"] - ref)
    assert np.all(err <= TOL * np.abs(ref) + 2e-10 * kf)


def test_jacobian(case):
    fop = case["fop"]
    fop.createJacobian(case["model"])
    J = fop.jacobian().numpy()
    ref = case["g"]["J"]
    assert J.shape == ref.shape
    assert _relmax(J, ref) < TOL
    # row-wise too: every data row within tolerance of its own scale
    rs = np.max(np.abs(ref), axis=1)
    assert np.max(np.max(np.abs(J - ref), axis=1) / rs) < TOL


def test_jacobian_operator(case):
    fop = case["fop"]
    fop.createJacobian(case["model"])
    Jop = fop.jacobian()
    J = Jop.numpy()
    rng = np.random.default_rng(5)
    x = rng.standard_normal(J.shape[1])
    y = rng.standard_normal(J.shape[0])
    assert _relmax(Jop.mult(x), J @ x) < 1e-12
    assert _relmax(Jop.transMult(y), J.T @ y) < 1e-12


def test_jacobian_homogeneous_analytic_branch(case):
    """no cached potentials + homogeneous model -> analytic potentials scaled by model[0]"""
    fop = _fop(case["mesh"], case["scheme"])
    hom = np.full(case["model"].size, 100.0)
    fop.createJacobian(hom)
    J = fop.jacobian().numpy()
    assert _relmax(J, case["g"]["J_hom"]) < TOL
    fop._core.close()


def test_against_compiled_reference_live(case):
    """same comparison against the reference run live on this box (different model seed)"""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(99)
    model = 10.0 ** (1.5 + 0.4 * rng.standard_normal(case["model"].size))
    R = ref.RefERT(case["mesh"], case["scheme"], sr=True)
    r_ref = R.response(model)
    J_ref = R.create_jacobian(model)
    fop = case["fop"]
    rhoa = fop.response(model)
    fop.createJacobian(model)
    J = fop.jacobian().numpy()
    kf = np.abs(case["scheme"].k)
    assert np.all(np.abs(rhoa - r_ref) <= TOL * np.abs(r_ref) + 2e-10 * kf)
    assert _relmax(J, J_ref) < TOL
    R.close()


# ---------------------------------------------------------------------------------------------
# variants of the path checked against the reference run live (oracle/_ref travels with the repo)
# ---------------------------------------------------------------------------------------------
def _live(mesh, scheme, model, sr=True, k=None, w=None):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    R = ref.RefERT(mesh, scheme, sr=sr)
    if k is not None:
        R.set_kw(k, w)
    r_ref = R.response(model)
    pots = R.subpotentials()
    J_ref = R.create_jacobian(model)
    R.close()
    return r_ref, pots, J_ref


def _ours(mesh, scheme, model, sr=True, k=None, w=None):
    fop = _fop(mesh, scheme, sr=sr)
    if k is not None:
        fop._core.setkValues(k)
        fop._core.setWeights(w)
    rhoa = fop.response(model)
    P = fop._core._plan
    pots = fop._core.get("pots").reshape(P.nS, P.N)
    fop.createJacobian(model)
    J = fop.jacobian().numpy()
    fop._core.close()
    return rhoa, pots, J


def _check(ours, refv, scheme):
    rhoa, pots, J = ours
    r_ref, p_ref, J_ref = refv
    kf = np.abs(scheme.k)
    assert np.all(np.abs(rhoa - r_ref) <= TOL * np.abs(r_ref) + 2e-10 * kf)
    for s in range(0, pots.shape[0], max(1, pots.shape[0] // 7)):
        assert _relmax(pots[s], p_ref[s]) < TOL
    assert _relmax(J, J_ref) < TOL


@pytest.mark.parametrize("name", ["2d_p1", "3d_p1"])
def test_total_field_variant_sr_false(name):
    """DCMultiElectrodeModelling (no singularity removal, dcfemmodelling.cpp:1755-1928)"""
    mesh, scheme, model = make_case(name)
    _check(_ours(mesh, scheme, model, sr=False), _live(mesh, scheme, model, sr=False), scheme)


@pytest.mark.parametrize("name", ["2d_p1", "2d_p2", "3d_p1"])
def test_free_electrodes(name):
    """sensors that do not coincide with mesh nodes -> ElectrodeShapeEntity (electrode.cpp:199-287)"""
    mesh, scheme, model = make_case(name)
    mesh.node_marker[:] = 0                       # no electrode nodes: every sensor is located in a cell
    sch = scheme.subset(np.arange(scheme.size))
    sch.sensors = scheme.sensors.copy()
    sch.sensors[:, 0] += 0.13                     # off the nodes, still on the surface
    from pygimli_b200.scheme import geometric_factors
    sch.k = geometric_factors(sch, mesh.dim)
    _check(_ours(mesh, sch, model), _live(mesh, sch, model), sch)


def test_buried_electrodes_crosshole():
    """electrodes below the surface: analytic primary potentials with a real mirror source (bertMisc.cpp:196-214)"""
    mesh, scheme, model = make_case("3d_crosshole")
    _check(_ours(mesh, scheme, model), _live(mesh, scheme, model), scheme)


def test_user_wavenumbers():
    """setkValues / setWeights override the default list (dcfemmodelling.h:239-243)"""
    mesh, scheme, model = make_case("2d_p1")
    k = np.array([0.01, 0.05, 0.2, 0.8, 2.5])
    w = np.array([0.02, 0.05, 0.2, 0.6, 1.1])
    _check(_ours(mesh, scheme, model, k=k, w=w), _live(mesh, scheme, model, k=k, w=w), scheme)


def test_dirichlet_boundary_marker():
    """faces with marker -3 become homogeneous Dirichlet rows/columns (dcfemmodelling.cpp:141-161)"""
    mesh, scheme, model = make_case("3d_p1")
    zb = mesh.pos[mesh.bounds].mean(1)[:, 2]
    mesh.bound_marker[np.isclose(zb, mesh.pos[:, 2].min())] = -3        # bottom of the box
    _check(_ours(mesh, scheme, model), _live(mesh, scheme, model), scheme)


def test_error_behaviour():
    from pygimli_b200 import _capi
    mesh, scheme, model = make_case("2d_p1")
    fop = _fop(mesh, scheme)
    bad = model.copy()
    bad[3] = -1.0
    with pytest.raises(_capi.PGB200Error, match="negative or zero resistivity"):
        fop.response(bad)
    with pytest.raises(_capi.PGB200Error, match="model length"):
        fop.response(model[:-1])
    fop._core.close()


def test_response_is_homogeneous_of_degree_one():
    """size-independent property: response(c * model) == c * response(model) (S scales with 1/c, the
    secondary-field right-hand side with it), and J(c * model) == J(model) for the same reason."""
    from pygimli_b200.mesh import graded_axis, grid_mesh_2d, mark_electrode_nodes
    from pygimli_b200.scheme import create_dd, geometric_factors
    ne = 9
    xs = graded_axis(0.0, 8.0, 0.5, 1.4, 40.0)
    ys = -graded_axis(0.0, 3.0, 0.5, 1.4, 40.0, both=False)
    mesh = grid_mesh_2d(xs, ys)                  # every quad is a model cell
    sens = np.zeros((ne, 3)); sens[:, 0] = np.arange(ne)
    mark_electrode_nodes(mesh, sens)
    sch = create_dd(sens); sch.k = geometric_factors(sch, 2)
    M = int(mesh.cell_marker.max()) + 1
    rng = np.random.default_rng(2)
    model = 10.0 ** (2 + 0.3 * rng.standard_normal(M))
    fop = _fop(mesh, sch)
    rhoa = fop.response(model)
    # scaling property: response(c * model) == c * response(model) to solver accuracy
    fop.createJacobian(model)
    J1 = fop.jacobian().numpy()
    r2 = fop.response(3.0 * model)
    assert np.max(np.abs(r2 - 3.0 * rhoa) / rhoa) < 1e-7
    fop.createJacobian(3.0 * model)
    J2 = fop.jacobian().numpy()
    assert np.max(np.abs(J2 - J1)) / np.max(np.abs(J1)) < 1e-7
    fop._core.close()
