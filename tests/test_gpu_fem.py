"""GPU parity of the generic FEM matrices on the path's element kernels:
CoreB200.fillStiffnessMatrix / fillMassMatrix against SparseMatrix::fillStiffnessMatrix / fillMassMatrix
(core/src/sparsematrix.h:1034-1065).  Checker: tests/golden/fem.npz, produced by the compiled reference
(tests/make_golden_fem.py); pattern bit-exact, values 1e-12 relative to the matrix scale (summation order differs)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from cases import FEM_CASES, make_case, fem_inputs

pytestmark = pytest.mark.gpu

GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD = np.load(os.path.join(GOLD_DIR, "fem.npz"))


@pytest.mark.parametrize("name", FEM_CASES)
def test_fill_matrices(name):
    from pygimli_b200 import CoreB200
    mesh, scheme, _ = make_case(name)
    a, b, X = fem_inputs(mesh)
    core = CoreB200(sr=True)
    core.setMesh(mesh)
    core.setData(scheme)
    g = np.load(os.path.join(GOLD_DIR, name + ".npz"))
    for tag, (rp, ci, vals) in (("K", core.fillStiffnessMatrix(a)), ("M", core.fillMassMatrix(b))):
        assert np.array_equal(rp, g["rowptr"]) and np.array_equal(ci, g["colidx"])
        S = sp.csr_matrix((vals, ci, rp), shape=(mesh.node_count, mesh.node_count))
        ref_x, ref_d = GOLD[f"{name}_{tag}x"], GOLD[f"{name}_{tag}diag"]
        assert np.max(np.abs(S @ X - ref_x)) <= 1e-12 * np.max(np.abs(ref_x))
        assert np.max(np.abs(S.diagonal() - ref_d)) <= 1e-12 * np.max(np.abs(ref_d))
    # scalar coefficient = constant per cell; stiffness annihilates constants, the unit mass matrix sums to the volume
    rp, ci, vK = core.fillStiffnessMatrix()
    K = sp.csr_matrix((vK, ci, rp), shape=(mesh.node_count, mesh.node_count))
    assert np.max(np.abs(K @ np.ones(mesh.node_count))) < 1e-10 * np.max(np.abs(K.diagonal()))
    rp, ci, vM = core.fillMassMatrix()
    assert abs(vM.sum() - mesh.cell_sizes().sum()) < 1e-10 * mesh.cell_sizes().sum()
    core.close()
