"""Pins the oracle's coverageDCtrans / createCoverage restatement (oracle/ert_oracle.py) to the reference's own
outputs: tests/golden/coverage.npz was produced by the compiled reference (tests/make_golden_coverage.py).  CPU-only."""
import os

import numpy as np
import pytest

from cases import coverage_case
from oracle.ert_oracle import coverage_dc_trans, create_coverage

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "coverage.npz"))


@pytest.mark.parametrize("dim", [2, 3])
def test_coverage_trans_matches_reference(dim):
    mesh, J, dd, mm, resp, model = coverage_case(dim)
    np.testing.assert_allclose(coverage_dc_trans(J, dd, mm), GOLD[f"cov_trans_{dim}d"], rtol=1e-14, atol=0.0)


@pytest.mark.parametrize("dim", [2, 3])
def test_create_coverage_matches_reference(dim):
    mesh, J, dd, mm, resp, model = coverage_case(dim)
    got = create_coverage(J, mesh.cell_marker, mesh.cell_sizes(), resp, model)
    np.testing.assert_allclose(got, GOLD[f"coverage_{dim}d"], rtol=1e-13, atol=0.0)
    unit = create_coverage(J, mesh.cell_marker, mesh.cell_sizes())
    np.testing.assert_allclose(unit, GOLD[f"coverage_unit_{dim}d"], rtol=1e-13, atol=0.0)


def test_create_coverage_size_mismatch_raises():
    mesh, J, dd, mm, resp, model = coverage_case(2)
    with pytest.raises(RuntimeError):
        create_coverage(J[:, :-1], mesh.cell_marker, mesh.cell_sizes(), resp, model[:-1])


def test_live_reference_when_present():
    """same check against the compiled reference run live (skipped on machines without oracle/_ref)"""
    from oracle import ref
    if not os.path.exists(ref.LIB_PATH):
        pytest.skip("oracle/_ref not built")
    mesh, J, dd, mm, resp, model = coverage_case(3)
    np.testing.assert_allclose(coverage_dc_trans(J, dd, mm), ref.coverage_trans(J, dd, mm), rtol=1e-14)
    np.testing.assert_allclose(create_coverage(J, mesh.cell_marker, mesh.cell_sizes(), resp, model),
                               ref.create_coverage(J, mesh, resp, model), rtol=1e-13)


def test_manager_coverage_host_part():
    """ERTManager.coverage() (ertManager.py:328-340) on top of coverageDCtrans: the host part (parameter sizes per marker,
    log10, per-cell look-up) with the oracle standing in for the GPU operator"""
    from pygimli_b200.ert_modelling import managerCoverage
    mesh, J, dd, mm, resp, model = coverage_case(2)

    class Stub:
        def coverageDCtrans(self, d, m):
            return coverage_dc_trans(J, d, m)
    got = managerCoverage(Stub(), mesh, resp, model)
    cov = coverage_dc_trans(J, 1.0 / resp, 1.0 / model)
    sizes = np.zeros(model.size)
    for c, mk in enumerate(mesh.cell_marker):                         # the reference's loop over paraDomain.cells()
        sizes[mk] += mesh.cell_sizes()[c]
    np.testing.assert_allclose(got, np.log10(cov / sizes)[mesh.cell_marker], rtol=1e-14)
    # several cells per marker: sizes add up
    mesh.cell_marker = (mesh.cell_marker // 2).astype(np.int32)
    M2 = int(mesh.cell_marker.max()) + 1
    got2 = managerCoverage(Stub2(J[:, :M2]), mesh, resp, model[:M2])
    sizes2 = np.bincount(mesh.cell_marker, weights=mesh.cell_sizes(), minlength=M2)
    np.testing.assert_allclose(got2, np.log10(coverage_dc_trans(J[:, :M2], 1.0 / resp, 1.0 / model[:M2]) / sizes2)[mesh.cell_marker], rtol=1e-14)


class Stub2:
    def __init__(self, J):
        self.J = J

    def coverageDCtrans(self, d, m):
        return coverage_dc_trans(self.J, d, m)
