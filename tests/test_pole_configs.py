"""Pole-dipole / dipole-pole / pole-pole rows (electrode index -1 = electrode at infinity) against the reference's own
outputs in tests/golden/pole_2d.npz (tests/make_golden_pole.py): analytic geometric factors (bertMisc.cpp:131-176) and
the oracle on the CPU; response and Jacobian through the C ABI on the GPU."""
import os

import numpy as np
import pytest

from cases import make_pole_case
from pygimli_b200.scheme import geometric_factors

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pole_2d.npz"))


def test_geometric_factors_with_poles():
    mesh, scheme, _ = make_pole_case()
    np.testing.assert_allclose(geometric_factors(scheme, 2), GOLD["kfac"], rtol=1e-13)


def test_oracle_with_poles():
    from oracle.ert_oracle import OracleERT
    mesh, scheme, model = make_pole_case()
    scheme.k = GOLD["kfac"]
    O = OracleERT(mesh, scheme)
    rhoa = O.response(model)
    assert np.max(np.abs(rhoa - GOLD["rhoa"]) / np.abs(GOLD["rhoa"])) < 1e-8
    J = O.jacobian(model)
    assert np.max(np.abs(J - GOLD["J"])) < 1e-9 * np.max(np.abs(GOLD["J"]))


@pytest.mark.gpu
def test_gpu_with_poles():
    from pygimli_b200 import ERTModellingB200
    mesh, scheme, model = make_pole_case()
    fop = ERTModellingB200(sr=True)
    fop.setMesh(mesh)
    fop.setData(scheme)                              # no k: response() fills the analytic factors (:1088-1093)
    rhoa = fop.response(model)
    np.testing.assert_allclose(fop._core._scheme.k, GOLD["kfac"], rtol=1e-13)
    assert np.all(np.abs(rhoa - GOLD["rhoa"]) <= 1e-8 * np.abs(GOLD["rhoa"]) + np.abs(GOLD["kfac"]) * 1e-10)
    fop.createJacobian(model)
    J = fop.jacobian().numpy()
    assert np.max(np.abs(J - GOLD["J"])) < 1e-8 * np.max(np.abs(GOLD["J"]))
