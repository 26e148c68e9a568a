"""The C-ABI library loads without a GPU and exports every symbol include/pgb200_ert.h declares;
host-only helpers work; compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from pygimli_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "pgb200_ert.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pgb200_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    names = _declared()
    assert len(names) >= 25
    L = _capi.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared but not exported: {missing}"
    assert set(_capi.EXPORTS) <= set(names)


def test_version_and_error_string():
    assert _capi.lib().pgb200_version() >= 100
    assert isinstance(_capi.last_error(), str)


def test_color_cells_is_conflict_free():
    from cases import make_case
    mesh, _, _ = make_case("3d_p2")
    color, n = _capi.color_cells(mesh.cells, mesh.node_count)
    assert n == color.max() + 1
    for c in range(n):
        nodes = mesh.cells[color == c].ravel()
        assert np.unique(nodes).size == nodes.size


def test_build_panels_covers_matrix():
    from cases import make_case
    from pygimli_b200.host_setup import build_pattern
    mesh, _, _ = make_case("3d_p1")
    rowptr, colidx, _ = build_pattern(mesh)
    pan = _capi.build_panels(rowptr, colidx, 64, 208)
    pp, hp = pan["panel_ptr"], pan["halo_ptr"]
    assert pp[0] == 0 and pp[-1] == mesh.node_count and np.all(np.diff(pp) > 0) and np.all(np.diff(pp) <= 64)
    assert np.all(np.diff(hp) <= 208)
    # local indices resolve to the original columns, self index to the diagonal
    rowof = np.repeat(np.arange(mesh.node_count), np.diff(rowptr))
    panel_of_row = np.repeat(np.arange(pan["n_panels"]), np.diff(pp))
    col_back = pan["halo_cols"][hp[panel_of_row[rowof]] + pan["lidx"][: colidx.size]]
    assert np.array_equal(col_back, colidx)
    self_back = pan["halo_cols"][hp[panel_of_row] + pan["self_idx"]]
    assert np.array_equal(self_back, np.arange(mesh.node_count))


def test_compute_fails_loudly_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    from cases import make_case
    from pygimli_b200 import ERTModellingB200
    mesh, scheme, model = make_case("2d_p1")
    fop = ERTModellingB200()
    fop.setMesh(mesh)
    fop.setData(scheme)
    with pytest.raises(_capi.PGB200Error, match="no CUDA device"):
        fop.response(model)
