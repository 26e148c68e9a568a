"""Groundwork for SURVEY.md §8(f).3 (complex resistivity): pins the oracle's restatement of the reference's complex
total-field path (oracle/ert_oracle.py: map_model_complex, total_field_complex, response_complex, jacobian_complex) to
the reference's own outputs in tests/golden/complex_2d_p1.npz (tests/make_golden_complex.py).  The CUDA side of this row
is not built yet; the 3-D golden file is there for it.  CPU-only."""
import os

import numpy as np
import pytest

from cases import complex_model, make_case
from oracle.ert_oracle import OracleERT

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def cplx():
    mesh, scheme, model = make_case("2d_p1")
    mc = complex_model(model)
    O = OracleERT(mesh, scheme)
    rhoa = O.response_complex(mc)
    return O, mc, rhoa, np.load(os.path.join(GOLD, "complex_2d_p1.npz"))


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def test_complex_response(cplx):
    O, mc, rhoa, g = cplx
    assert _rel(rhoa, g["rhoa"]) < 1e-10
    assert np.all(rhoa.imag < 0.0)                                    # negative phases in, negative phases out


def test_complex_potentials(cplx):
    O, mc, rhoa, g = cplx
    sol = sum(O.w[kk] * O.pots_c[kk * O.nE:(kk + 1) * O.nE] for kk in range(len(O.k)))
    assert _rel(sol[[0, O.nE - 1]], g["sol_rows"]) < 1e-10


def test_complex_jacobian(cplx):
    O, mc, rhoa, g = cplx
    J = O.jacobian_complex(mc, O.pots_c)
    assert _rel(J, g["J"]) < 1e-10


def test_vanishing_phase_reduces_to_the_real_total_field_path(cplx):
    """phases of 1e-9 rad: the real part of the complex solve is the real total-field solve (an exactly zero imaginary
    part is not a valid input: the reference prolongates |values| < 1e-12 as 'empty' cells, :1203-1206)"""
    O, mc, rhoa, g = cplx
    real_model = np.abs(mc)
    pots_c = O.total_field_complex(real_model * np.exp(-1e-9j))
    pots_r = O.total_field(real_model)
    assert np.max(np.abs(pots_c.imag)) <= 1e-8 * np.max(np.abs(pots_r))
    assert _rel(pots_c.real, pots_r) < 1e-10
