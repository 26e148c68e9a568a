"""Pins the oracle's generic FEM matrices (OracleERT.fill_matrix: fillStiffnessMatrix / fillMassMatrix,
core/src/sparsematrix.h:1034-1065) to the reference's outputs in tests/golden/fem.npz (tests/make_golden_fem.py).
CPU-only; the 2-D cases (the 3-D ones take the pure-Python element loop too long and are covered on the GPU side)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from cases import make_case, fem_inputs
from oracle.ert_oracle import OracleERT

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fem.npz"))


@pytest.mark.parametrize("name", ["2d_p1", "2d_p2"])
def test_fill_matrices_match_reference(name):
    mesh, scheme, _ = make_case(name)
    a, b, X = fem_inputs(mesh)
    O = OracleERT(mesh, scheme)
    rp, ci = O.pattern()
    for tag, kw in (("K", dict(a=a)), ("M", dict(b=b))):
        S = sp.csr_matrix((O.fill_matrix(**kw), ci, rp), shape=(O.N, O.N))
        ref_x, ref_d = GOLD[f"{name}_{tag}x"], GOLD[f"{name}_{tag}diag"]
        assert np.max(np.abs(S @ X - ref_x)) <= 1e-12 * np.max(np.abs(ref_x))
        assert np.max(np.abs(S.diagonal() - ref_d)) <= 1e-12 * np.max(np.abs(ref_d))


def test_stiffness_annihilates_constants_and_mass_sums_to_volume():
    mesh, scheme, _ = make_case("2d_p2")
    O = OracleERT(mesh, scheme)
    rp, ci = O.pattern()
    K = sp.csr_matrix((O.fill_matrix(a=np.ones(O.C)), ci, rp), shape=(O.N, O.N))
    M = sp.csr_matrix((O.fill_matrix(b=np.ones(O.C)), ci, rp), shape=(O.N, O.N))
    assert np.max(np.abs(K @ np.ones(O.N))) < 1e-10 * np.max(np.abs(K.diagonal()))
    assert abs(M.sum() - mesh.cell_sizes().sum()) < 1e-10 * mesh.cell_sizes().sum()
