"""Single-GPU coverage of what one rank of a multi-GPU run executes (the driver's GPU box has one GPU, so
tests/test_gpu_multi.py is skipped there): a source shard [c0, c1) solved alone must reproduce the columns of the full
solve, and shards of at most 32 columns must take the gather-form SpMM (k_spmm_gather), wider ones the
partial-width tiles of the streamed kernel."""
import numpy as np
import pytest

from cases import make_wide_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full():
    from pygimli_b200.dist import ShardedERT
    mesh, scheme, model, rows = make_wide_case("3d_p1_wide")
    fop = ShardedERT(mesh, scheme)
    fop.response(model)
    P = fop.core._plan
    pots = fop.core.get("pots").reshape(P.nS, P.N).copy()
    yield dict(mesh=mesh, scheme=scheme, model=model, pots=pots, nS=P.nS, N=P.N, D=fop.D)
    fop.core.close()


@pytest.mark.parametrize("window,path", [((0, 9), "gather"), ((30, 47), "gather"), ((63, 72), "gather"), ((41, 72), "gather"),
                                          ((0, 36), "staged"), ((35, 72), "staged")])
def test_source_shard_reproduces_full_solve(full, window, path):
    from pygimli_b200.dist import ShardedERT
    a, b = window
    f2 = ShardedERT(full["mesh"], full["scheme"])
    f2.core.setShard(a, b, 0, full["D"] * (b - a) // full["nS"])
    f2.response(full["model"])
    info = f2.core.pathInfo()
    pots = f2.core.get("pots").reshape(full["nS"], full["N"])
    ref = full["pots"][a:b]
    assert np.max(np.abs(pots[a:b] - ref)) <= 1e-9 * np.max(np.abs(ref))
    if path == "gather":
        assert info["spmm_panel_nc"] >= 200                 # 200 + n-tiles: k_spmm_gather
    else:
        assert 100 <= info["spmm_panel_nc"] < 200           # 100 + n-tiles: k_spmm_mma on a partial-width tile
    st = f2.core.stats()
    assert st["max_rel_residual"] <= 1.0e-12 * 1.0001
    f2.core.close()
