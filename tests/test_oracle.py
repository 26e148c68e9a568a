"""Pins oracle/ert_oracle.py (the CPU restatement) to the reference's own outputs:
golden vectors in tests/golden/*.npz were produced by the compiled reference (tests/make_golden.py).
CPU-only; small cases so the whole file runs in well under a minute."""
import os

import numpy as np
import pytest

from cases import make_case
from oracle.ert_oracle import OracleERT, bessel_k0, bessel_k1, kwave_list

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.fixture(scope="module", params=["2d_p1", "2d_p2"])
def small(request):
    mesh, scheme, model = make_case(request.param)
    return request.param, mesh, scheme, model, np.load(os.path.join(GOLD, request.param + ".npz")), OracleERT(mesh, scheme)


def test_pattern_bit_exact(small):
    _, _, _, _, g, O = small
    rp, ci = O.pattern()
    assert np.array_equal(rp, g["rowptr"]) and np.array_equal(ci, g["colidx"])


def test_wavenumbers_bit_exact(small):
    _, _, _, _, g, O = small
    assert np.array_equal(O.k, g["k"]) and np.array_equal(O.w, g["w"])


def test_electrodes(small):
    _, _, _, _, g, O = small
    assert np.array_equal(np.asarray(O.el_node), g["el_nodes"])


def test_model_mapping(small):
    _, _, _, model, g, O = small
    assert _rel(O.map_model(model), g["rho"]) < 1e-13


def test_matrix_values(small):
    _, _, _, _, g, O = small
    for k, key in ((g["k"][0], "vals_k0"), (g["k"][-1], "vals_klast")):
        v = O.matrix_values(float(k), g["rho"])
        assert np.max(np.abs(v - g[key])) / np.max(np.abs(g[key])) < 1e-12


def test_response_and_potentials_and_jacobian(small):
    name, _, scheme, model, g, O = small
    rhoa = O.response(model)
    assert np.all(np.abs(rhoa - g["rhoa"]) <= 1e-8 * np.abs(g["rhoa"]) + 2e-10 * np.abs(scheme.k))
    for r, ref in zip(g["pots_rows"], g["pots"]):
        assert _rel(O.pots[r], ref) < 1e-8
    J = O.jacobian(model)
    assert _rel(J, g["J"]) < 1e-8


def test_jacobian_analytic_branch(small):
    _, mesh, scheme, model, g, _ = small
    O = OracleERT(mesh, scheme)
    J = O.jacobian(np.full(model.size, 100.0))
    assert _rel(J, g["J_hom"]) < 1e-8


def test_3d_matrix_and_mapping():
    mesh, scheme, model = make_case("3d_p2")
    g = np.load(os.path.join(GOLD, "3d_p2.npz"))
    O = OracleERT(mesh, scheme)
    assert np.array_equal(O.k, g["k"])
    rho = O.map_model(model)
    assert _rel(rho, g["rho"]) < 1e-13
    v = O.matrix_values(0.0, g["rho"])
    assert np.max(np.abs(v - g["vals_k0"])) / np.max(np.abs(g["vals_k0"])) < 1e-12


def test_live_reference_if_built():
    """Bessel and wavenumber routines against the compiled reference itself (when oracle/_ref exists)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    x = np.logspace(-5, 2, 500)
    k0, k1 = ref.bessel(x)
    assert np.max(np.abs(np.array([bessel_k0(v) for v in x]) - k0) / np.abs(k0)) < 1e-14
    assert np.max(np.abs(np.array([bessel_k1(v) for v in x]) - k1) / np.abs(k1)) < 1e-14
    sens = np.zeros((7, 3))
    sens[:, 0] = np.arange(7) * 1.5
    k, w = kwave_list(2, sens)
    kr, wr = ref.kwave_list(0.75, 18.0, max(int(np.floor(6 * np.log10(18.0 / 0.75))), 4), 4)
    assert np.array_equal(k, kr) and np.array_equal(w, wr)
