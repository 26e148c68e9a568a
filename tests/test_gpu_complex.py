"""GPU parity of the complex-resistivity path: CoreB200.setComplex(True) against golden vectors the
reference produced with DCMultiElectrodeModelling::setComplex(true) (tests/make_golden_complex.py ->
tests/golden/complex_*.npz; core/src/bert/dcfemmodelling.cpp:235-242, 1103-1118, 1410-1461, 1755-1925).
Tolerance (north_star): apparent resistivities and Jacobian 1e-8 relative."""
import os

import numpy as np
import pytest

from cases import COMPLEX_CASES, complex_model, make_case

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-8


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.fixture(scope="module", params=COMPLEX_CASES)
def cx(request):
    from pygimli_b200 import ERTModellingB200
    name = request.param
    mesh, scheme, model = make_case(name)
    mc = complex_model(model)
    g = np.load(os.path.join(GOLD, "complex_" + name + ".npz"))
    fop = ERTModellingB200(sr=False)
    fop.setComplex(True)
    fop.setMesh(mesh)
    fop.setData(scheme)
    m2 = np.concatenate([mc.real, mc.imag])
    resp = fop.response(m2)
    yield dict(name=name, fop=fop, g=g, mc=mc, m2=m2, resp=resp, D=scheme.size)
    fop._core.close()


def test_complex_response(cx):
    D = cx["D"]
    rhoa = cx["resp"][:D] + 1j * cx["resp"][D:]
    assert _rel(rhoa, cx["g"]["rhoa"]) < TOL
    assert np.all(rhoa.imag < 0.0)


def test_complex_jacobian(cx):
    fop = cx["fop"]
    Jsq = fop.createJacobian(cx["m2"])
    J = fop._core.jacobian().numpy()
    ref = cx["g"]["J"]
    assert J.shape == ref.shape
    assert _rel(J, ref) < TOL
    rs = np.max(np.abs(ref), axis=1)
    assert np.max(np.max(np.abs(J - ref), axis=1) / rs) < TOL
    D, M = J.shape
    assert Jsq.shape == (2 * D, 2 * M)                       # pg.utils.squeezeComplex layout (ertModelling.py:231-235)
    assert np.array_equal(Jsq[:D, :M], J.real) and np.array_equal(Jsq[:D, M:], -J.imag)
    assert np.array_equal(Jsq[D:, :M], J.imag) and np.array_equal(Jsq[D:, M:], J.real)


def test_complex_needs_the_total_field_operator():
    from pygimli_b200 import ERTModellingB200, _capi
    fop = ERTModellingB200(sr=True)
    with pytest.raises(_capi.PGB200Error, match="total-field"):
        fop.setComplex(True)


def test_vanishing_phase_reduces_to_the_real_path(cx):
    """phases of 1e-9 rad: the real part equals the real total-field response (no rounding / reciprocity mean in the complex path)"""
    from pygimli_b200 import ERTModellingB200
    mesh, scheme, model = make_case(cx["name"])
    fr = ERTModellingB200(sr=False)
    fr.setMesh(mesh); fr.setData(scheme)
    real = fr.response(model)
    fr._core.close()
    mc = model * np.exp(-1e-9j)
    resp = cx["fop"].response(np.concatenate([mc.real, mc.imag]))
    D = cx["D"]
    assert np.max(np.abs(resp[:D] - real) / np.abs(real)) < 1e-6       # the real path rounds u to 1e-10 and averages reciprocals
    assert np.max(np.abs(resp[D:])) <= 1e-7 * np.max(np.abs(real))
