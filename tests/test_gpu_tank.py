"""GPU parity on a closed tank: pure-Neumann 3-D domain with calibration nodes
(-1000, or the reference's node 0) as homogeneous Dirichlet rows and a reference-electrode node (-999) as the sink of
every current pattern (core/src/bert/dcfemmodelling.cpp:141-161, 1009-1064, 1517-1523, 1868-1870), total-field operator,
numeric geometric factors (:1539-1556).  Golden vectors: tests/make_golden_tank.py (the compiled reference)."""
import os

import numpy as np
import pytest

from cases import make_tank_case

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-8
CASES = {"tank_ref_cal": (True, True), "tank_ref": (True, False), "tank_last": (False, True)}


def _relmax(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.fixture(scope="module", params=list(CASES))
def tank(request):
    from pygimli_b200 import ERTModellingB200
    name = request.param
    mesh, scheme, model = make_tank_case(*CASES[name])
    g = np.load(os.path.join(GOLD, name + ".npz"))
    fop = ERTModellingB200(sr=False)
    fop.setMesh(mesh)
    fop.setData(scheme)
    k = fop._core.calcGeometricFactor()
    fop._core.setGeometricFactors(k)
    rhoa = fop.response(model)
    yield dict(name=name, fop=fop, g=g, model=model, k=k, rhoa=rhoa, mesh=mesh)
    fop._core.close()


def test_plan_flags(tank):
    P = tank["fop"]._core._plan
    assert P.topography and P.neumann_domain                      # Neumann domain => topography (:750-755)
    assert (P.ref_node >= 0) == CASES[tank["name"]][0] and bool(P.ref_last) == (not CASES[tank["name"]][0])
    assert len(P.dir_nodes) == 1                                  # the calibration node (or the reference's node 0)


def test_numeric_geometric_factors(tank):
    assert _relmax(tank["k"], tank["g"]["kfac"]) < TOL


def test_tank_potentials(tank):
    core = tank["fop"]._core
    P = core._plan
    pots = core.get("pots").reshape(P.nS, P.N)
    ref = tank["g"]["pots"]
    for r in range(ref.shape[0]):
        assert _relmax(pots[r], ref[r]) < TOL
    if ref.shape[0] < P.nS:                                       # last electrode = current reference: no pattern of its own
        assert not np.any(pots[ref.shape[0]:])


def test_tank_apparent_resistivity(tank):
    ref = tank["g"]["rhoa"]
    assert np.all(np.abs(tank["rhoa"] - ref) <= TOL * np.abs(ref) + 2e-10 * np.abs(tank["k"]))


def test_tank_jacobian(tank):
    from pygimli_b200 import _capi
    fop = tank["fop"]
    if "J" not in tank["g"].files:
        with pytest.raises(_capi.PGB200Error, match="rowsize to small"):       # the reference throws the same length error
            fop.createJacobian(tank["model"])
        return
    fop.createJacobian(tank["model"])
    J = fop.jacobian().numpy()
    ref = tank["g"]["J"]
    assert J.shape == ref.shape
    rs = np.max(np.abs(ref), axis=1)
    assert np.max(np.max(np.abs(J - ref), axis=1) / rs) < TOL


def test_singularity_removal_refuses_reference_electrodes():
    from pygimli_b200 import ERTModellingB200, _capi
    mesh, scheme, model = make_tank_case(True, True)
    scheme.k = np.ones(scheme.size)
    fop = ERTModellingB200(sr=True)
    fop.setMesh(mesh)
    fop.setData(scheme)
    with pytest.raises(Exception, match="reference"):
        fop.response(model)
    fop._core.close()
