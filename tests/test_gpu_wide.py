"""GPU parity for the kernels bench.py actually times ("what's weak" 1-2), against golden vectors the
reference produced (tests/make_golden_wide.py -> tests/golden/wide_*.npz):

  3d_p1_wide  72 sources: two column tiles of the panel-staged SpMM in its three epilogue roles, replayed as CUDA graphs,
              and a 72 x 72 Gram block in the Jacobian
  2d_p1_wide  65 electrodes x 18 wavenumbers: column tiles that straddle wavenumber groups (the two_k path)
  3d_p1_192   192 sources: the current-electrode list needs several Jacobian chunks

Tolerance (north_star): potentials, apparent resistivities and Jacobian 1e-8 relative; rhoa additionally one quantum of
the reference's round(u, 1e-10) * |k|.  Every test also asserts that the intended kernel path ran (pathInfo)."""
import os

import numpy as np
import pytest

from cases import WIDE_CASES, make_wide_case

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-8


def _relmax(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module", params=WIDE_CASES)
def wide(request):
    from pygimli_b200 import ERTModellingB200
    name = request.param
    mesh, scheme, model, rows = make_wide_case(name)
    g = np.load(os.path.join(GOLD, "wide_" + name + ".npz"))
    fop = ERTModellingB200(sr=True)
    fop.setMesh(mesh)
    fop.setData(scheme)
    rhoa = fop.response(model)
    info = fop._core.pathInfo()
    yield dict(name=name, mesh=mesh, scheme=scheme, model=model, rows=rows, g=g, fop=fop, rhoa=rhoa, info=info)
    fop._core.close()


def test_intended_solver_path_ran(wide):
    info, name = wide["info"], wide["name"]
    assert info["amg_levels"] >= 2                      # multilevel preconditioner active
    assert info["spmm_panel_nc"] >= 1                   # panel-staged SpMM, not the plain gather kernel
    assert info["graph_launches"] >= 1                  # PCG iterations replayed as CUDA graphs
    assert info["spmm_slots"] >= 2 and info["stream_levels"] >= 0
    if name == "3d_p1_wide":
        assert info["spmm_panel_nc"] >= 2               # 72 columns: two column pairs per lane
    if name == "2d_p1_wide":
        assert info["spmm_tiles"] >= 2                  # one column tile per wavenumber group
    st = wide["fop"]._core.stats()
    assert st["max_rel_residual"] <= 1.0e-12 * 1.0001


def test_wide_potentials(wide):
    core = wide["fop"]._core
    P = core._plan
    assert np.array_equal(P.k, wide["g"]["k"]) and np.array_equal(P.w, wide["g"]["w"])
    pots = core.get("pots").reshape(P.nS, P.N)
    for r, ref in zip(wide["g"]["pots_rows"], wide["g"]["pots"]):
        assert _relmax(pots[r], ref) < TOL


def test_wide_apparent_resistivity(wide):
    ref = wide["g"]["rhoa"]
    kf = np.abs(wide["g"]["kfac"])
    assert np.all(np.abs(wide["rhoa"] - ref) <= TOL * np.abs(ref) + 2e-10 * kf)


def test_wide_jacobian_rows(wide):
    fop = wide["fop"]
    fop.createJacobian(wide["model"])
    info = fop._core.pathInfo()
    if wide["name"] == "3d_p1_192":
        assert info["jac_chunks"] >= 2 or info["jac_tiles_per_thread"] >= 2
    J = fop.jacobian().numpy()[wide["rows"]]
    ref = wide["g"]["J"]
    assert J.shape == ref.shape
    assert _relmax(J, ref) < TOL
    rs = np.max(np.abs(ref), axis=1)
    assert np.max(np.max(np.abs(J - ref), axis=1) / rs) < TOL
