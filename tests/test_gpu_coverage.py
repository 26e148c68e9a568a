"""GPU parity of the inversion-side Jacobian operators: the error-/transform-weighted products
(MultLeftRightMatrix, pygimli/frameworks/inversion.py:705-708) and the coverage (coverageDCtrans / createCoverage,
core/src/bert/bertJacobian.cpp:569-628), all on the HBM-resident J through the C ABI.

Checker: the oracle restatement (pinned to the reference by tests/test_coverage_oracle.py) and, when oracle/_ref is
present, the compiled reference itself on the same matrix.  Tolerance: 1e-12 relative (sums of D non-negative terms /
dot products of length D or M in a different order, measured against the size of the summed terms)."""
import os

import numpy as np
import pytest

from cases import make_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["2d_p1", "3d_p1"])
def jac(request):
    from pygimli_b200 import ERTModellingB200
    mesh, scheme, _ = make_case(request.param)
    # one model entry per cell and no background region: the forward mesh is its own parameter mesh
    mesh.cell_marker = np.random.default_rng(5).permutation(mesh.cell_count).astype(np.int32)
    model = 10.0 ** (2.0 + 0.5 * np.random.default_rng(6).standard_normal(mesh.cell_count))
    fop = ERTModellingB200(sr=True)
    fop.setMesh(mesh)
    fop.setData(scheme)
    resp = fop.response(model)
    fop.createJacobian(model)
    J = fop.jacobian()
    return dict(mesh=mesh, model=model, resp=resp, J=J, Jn=J.numpy(), fop=fop)


def _close(a, b, tol=1e-12, scale=None):
    """|a - b| <= tol * scale; scale defaults to max|b|, dot products pass the size of their summed terms"""
    assert a.shape == b.shape
    assert np.max(np.abs(a - b)) <= tol * (np.max(np.abs(b)) if scale is None else scale)


def test_mult_lr(jac):
    from pygimli_b200.ert_modelling import MultLeftRightMatrixB200
    rng = np.random.default_rng(1)
    J, Jn = jac["J"], jac["Jn"]
    left, right = rng.standard_normal(J.rows()), rng.standard_normal(J.cols())
    x, y = rng.standard_normal(J.cols()), rng.standard_normal(J.rows())
    A = MultLeftRightMatrixB200(J, left, right)
    Ja = np.abs(Jn)
    _close(A.mult(x), left * (Jn @ (right * x)), scale=np.max(np.abs(left) * (Ja @ np.abs(right * x))))
    _close(A.transMult(y), right * (Jn.T @ (left * y)), scale=np.max(np.abs(right) * (Ja.T @ np.abs(left * y))))
    _close(J.mult_lr(x, left=left), left * (Jn @ x), scale=np.max(np.abs(left) * (Ja @ np.abs(x))))
    _close(J.transMult_lr(y, right=right), right * (Jn.T @ y), scale=np.max(np.abs(right) * (Ja.T @ np.abs(y))))
    _close(J.mult_lr(x), J.mult(x), scale=np.max(Ja @ np.abs(x)))      # same kernel; the column-chunk atomics may reorder
    with pytest.raises(Exception):
        MultLeftRightMatrixB200(J, left[:-1], right)
    with pytest.raises(ValueError):
        J.mult_lr(x[:-1])


def test_adjoint_identity(jac):
    """<A x, y> == <x, A^T y> for the weighted operator"""
    from pygimli_b200.ert_modelling import MultLeftRightMatrixB200
    rng = np.random.default_rng(2)
    J = jac["J"]
    A = MultLeftRightMatrixB200(J, 1.0 / jac["resp"], jac["model"])
    x, y = rng.standard_normal(J.cols()), rng.standard_normal(J.rows())
    a, b = float(A.mult(x) @ y), float(x @ A.transMult(y))
    assert abs(a - b) <= 1e-10 * max(abs(a), abs(b))


def test_coverage_trans(jac):
    from oracle.ert_oracle import coverage_dc_trans
    from pygimli_b200.ert_modelling import coverageDCtrans
    J, Jn = jac["J"], jac["Jn"]
    dd, mm = 1.0 / jac["resp"], 1.0 / jac["model"]
    got = coverageDCtrans(J, dd, mm)
    _close(got, coverage_dc_trans(Jn, dd, mm))
    _close(J.coverageDCtrans(dd) / np.abs(mm), got)                    # undivided partial sums (row shards)
    assert np.all(got >= 0.0)


def test_create_coverage(jac):
    from oracle import ref
    from oracle.ert_oracle import create_coverage
    from pygimli_b200.ert_modelling import createCoverage
    J, Jn, mesh = jac["J"], jac["Jn"], jac["mesh"]
    got = createCoverage(J, mesh, jac["resp"], jac["model"])
    _close(got, create_coverage(Jn, mesh.cell_marker, mesh.cell_sizes(), jac["resp"], jac["model"]))
    _close(createCoverage(J, mesh), create_coverage(Jn, mesh.cell_marker, mesh.cell_sizes()))
    if os.path.exists(ref.LIB_PATH):
        _close(got, ref.create_coverage(Jn, mesh, jac["resp"], jac["model"]))
    from pygimli_b200 import MeshArrays
    fewer = MeshArrays(mesh.dim, mesh.pos, mesh.node_marker, mesh.cells[:-1], mesh.cell_marker[:-1] % (J.cols() - 1),
                       mesh.bounds, mesh.bound_marker)
    with pytest.raises(RuntimeError):                                  # the reference logs "Coverage fails" here
        createCoverage(J, fewer, jac["resp"], jac["model"])
