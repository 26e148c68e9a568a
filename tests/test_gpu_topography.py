"""GPU parity of the topography branch: numeric primary potentials from a P2 total-field solve
on the GPU (checkPrimpotentials_, core/src/bert/dcfemmodelling.cpp:2009-2056), numeric geometric factors
(calcGeometricFactor, :1539-1556), singularity-removal response and Jacobian on a mesh with a hill.

Checker: golden vectors the compiled reference produced here (tests/golden/topo_*.npz, tests/make_golden_topo.py).
Tolerances as in test_gpu_parity.py: 1e-8 relative on geometric factors, potentials, rhoa (plus the round(u, 1e-10)
quantum) and J."""
import os

import numpy as np
import pytest

from cases import TOPO_CASES, make_topo_case

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-8


def _relmax(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.fixture(scope="module", params=TOPO_CASES)
def topo(request):
    from pygimli_b200 import ERTModellingB200
    mesh, scheme, model = make_topo_case(request.param)
    g = np.load(os.path.join(GOLD, request.param + ".npz"))
    fop = ERTModellingB200(sr=True)
    fop.setMesh(mesh)
    fop.setData(scheme)
    return dict(name=request.param, mesh=mesh, scheme=scheme, model=model, g=g, fop=fop)


def test_response_needs_k_factors(topo):
    """response() on a topography mesh without k-factors throws in the reference (dcfemmodelling.cpp:1096-1098)"""
    core = topo["fop"]._core
    assert core._ensure_plan().topography
    with pytest.raises(RuntimeError, match="K-factors"):
        core.response(topo["model"])


def test_numeric_geometric_factors(topo):
    core = topo["fop"]._core
    k = core.calcGeometricFactor()
    assert _relmax(k, topo["g"]["kfac"]) < TOL
    assert core.primary_stats["max_rel_residual"] <= 1e-12


def test_primary_potentials(topo):
    core, g = topo["fop"]._core, topo["g"]
    core._ensure_handle()
    prim = core.get("prim").reshape(-1, core._plan.N)
    assert _relmax(prim[[0, prim.shape[0] - 1]], g["prim_rows"]) < TOL


def test_response_and_potentials(topo):
    core, g = topo["fop"]._core, topo["g"]
    core.setGeometricFactors(core.calcGeometricFactor())
    rhoa = core.response(topo["model"])
    quantum = np.abs(g["kfac"]) * 1e-10
    assert np.all(np.abs(rhoa - g["rhoa"]) <= TOL * np.abs(g["rhoa"]) + quantum)
    pots = core.get("pots").reshape(-1, core._plan.N)
    assert _relmax(pots[[0, pots.shape[0] - 1]], g["pot_rows"]) < TOL


def test_jacobian(topo):
    core, g = topo["fop"]._core, topo["g"]
    core.setGeometricFactors(g["kfac"])
    core.response(topo["model"])
    core.createJacobian(topo["model"])
    assert _relmax(core.jacobian().numpy(), g["J"]) < TOL


def test_homogeneous_model_is_solved_numerically(topo):
    """with topography createJacobian never takes the analytic branch (:1272) and rhoa == rho up to the k-factors"""
    core, g = topo["fop"]._core, topo["g"]
    core.setGeometricFactors(g["kfac"])
    hom = np.full(topo["model"].size, 100.0)
    rhoa = core.response(hom)
    assert np.all(np.abs(rhoa - g["rhoa_hom"]) <= TOL * 100.0 + np.abs(g["kfac"]) * 1e-10)
    core.clearPotentials()
    core.createJacobian(hom)
    assert core.stats()["pcg_iterations"] >= 0 and core.stats()["solves"] >= 1


def test_jacobian_without_k_uses_numeric_factors():
    """createJacobian on a fresh fop fills the missing k-factors numerically first (:1286-1290)"""
    from pygimli_b200 import ERTModellingB200
    mesh, scheme, model = make_topo_case("topo_2d")
    g = np.load(os.path.join(GOLD, "topo_2d.npz"))
    fop = ERTModellingB200(sr=True)
    fop.setMesh(mesh)
    fop.setData(scheme)
    fop.createJacobian(model)
    assert _relmax(fop._core._scheme.k, g["kfac"]) < TOL
    assert _relmax(fop.jacobian().numpy(), g["J"]) < TOL


def test_total_field_numeric_factors_without_k():
    """sr=False on a topography mesh without k-factors: calcGeometricFactor takes 1 / (u(rho = 1) + TOLERANCE) from a
    rho = 1 solve on this mesh (dcfemmodelling.cpp:1527-1556), and createJacobian fills the factors the same way
    (:1286-1290) instead of raising"""
    from oracle import ref
    from pygimli_b200 import ERTModellingB200
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    mesh, scheme, model = make_topo_case("topo_2d")
    R = ref.RefERT(mesh, scheme, sr=False)
    k_ref = R.geometric_factors()
    R.set_k(k_ref)
    r_ref = R.response(model)
    J_ref = R.create_jacobian(model)
    R.close()
    fop = ERTModellingB200(sr=False)
    fop.setMesh(mesh)
    fop.setData(scheme)
    k = fop._core.calcGeometricFactor()
    assert _relmax(k, k_ref) < TOL
    fop2 = ERTModellingB200(sr=False)
    fop2.setMesh(mesh)
    fop2.setData(scheme)
    fop2.createJacobian(model)                       # fills k numerically first
    assert _relmax(fop2._core._scheme.k, k_ref) < TOL
    rhoa = fop2.response(model)
    assert np.all(np.abs(rhoa - r_ref) <= TOL * np.abs(r_ref) + np.abs(k_ref) * 2e-10)
    assert _relmax(fop2.jacobian().numpy(), J_ref) < TOL
    fop._core.close()
    fop2._core.close()
