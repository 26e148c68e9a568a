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


def _replay_stream_spmm(pan, vals, X, tile):
    """CPU replay of k_spmm_stream's traversal: panels -> chunks (last to first) -> rows -> re-ordered entries, reading X
    only through the staged halo rows of the chunk and the tile's columns"""
    c0, c1 = tile
    Y = np.zeros_like(X)
    for p in range(pan["n_panels"]):
        r0, r1 = pan["panel_row_ptr"][p], pan["panel_row_ptr"][p + 1]
        acc = np.zeros((r1 - r0, c1 - c0))
        for ch in range(pan["panel_chunk_ptr"][p + 1] - 1, pan["panel_chunk_ptr"][p] - 1, -1):
            h0, h1 = pan["chunk_halo_ptr"][ch], pan["chunk_halo_ptr"][ch + 1]
            staged = X[pan["halo_cols"][h0:h1], c0:c1]                      # what the TMA copies bring
            e0 = pan["chunk_ent_ptr"][ch]
            crp = pan["crp"][ch * pan["crp_stride"]: (ch + 1) * pan["crp_stride"]]
            for r in range(r1 - r0):
                for e in range(e0 + crp[r], e0 + crp[r + 1]):
                    acc[r] += vals[pan["ent_src"][e]] * staged[pan["ent_idx"][e]]
            if ch == pan["panel_chunk_ptr"][p]:
                # chunk 0 starts with the panel's own rows
                assert np.array_equal(pan["halo_cols"][h0:h0 + (r1 - r0)], np.arange(r0, r1))
        Y[r0:r1, c0:c1] = acc
    return Y


@pytest.mark.parametrize("name,limits", [("3d_p1", (60, 104, 2)), ("3d_p2", (60, 104, 2)), ("2d_p1", (60, 104, 2)),
                                         ("3d_p1", (16, 40, 3)), ("2d_p2", (8, 30, 4))])
def test_stream_panels_reproduce_spmm(name, limits):
    """the streamed row-panel layout (csrc/stream_panels.h) visits every CSR entry exactly once and its staged-halo
    indices resolve to the right columns: a CPU replay of the kernel's traversal equals A @ X"""
    import scipy.sparse as sp
    from cases import make_case
    from pygimli_b200.host_setup import build_pattern
    mesh, _, _ = make_case(name)
    rowptr, colidx, _ = build_pattern(mesh)
    rmax, hc, nch = limits
    pan = _capi.build_stream_panels(rowptr, colidx, rmax, hc, nch)
    N = mesh.node_count
    pr = pan["panel_row_ptr"]
    assert pr[0] == 0 and pr[-1] == N and np.all(np.diff(pr) > 0) and np.all(np.diff(pr) <= rmax)
    assert np.all(np.diff(pan["chunk_halo_ptr"]) <= hc) and np.all(np.diff(pan["panel_chunk_ptr"]) <= nch)
    assert pan["crp_stride"] % 4 == 0 and pan["nnz"] == colidx.size
    assert np.array_equal(np.sort(pan["ent_src"][:colidx.size]), np.arange(colidx.size))       # a permutation of the CSR slots
    rng = np.random.default_rng(0)
    vals = rng.standard_normal(colidx.size)
    X = rng.standard_normal((N, 6))
    A = sp.csr_matrix((vals, colidx, rowptr), shape=(N, N))
    Y = _replay_stream_spmm(pan, vals, X, (1, 5))
    ref = np.zeros_like(X)
    ref[:, 1:5] = (A @ X)[:, 1:5]
    assert np.max(np.abs(Y - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_stream_panels_reject_too_wide_rows():
    rowptr = np.array([0, 5, 6, 7, 8, 9], np.int32)
    colidx = np.array([0, 1, 2, 3, 4, 1, 2, 3, 4], np.int32)
    with pytest.raises(_capi.PGB200Error, match="halo limit"):
        _capi.build_stream_panels(rowptr, colidx, 2, 2, 2)


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


def _replay_mma_spmm(pan, vals, X):
    """CPU replay of k_spmm_mma: per panel, chunks last to first; per (chunk, group) the k-steps: an 8 x 4 block of A
    values (fragment order: element 4 * row + column) times the 4 staged rows the k-step's word names"""
    N = X.shape[0]
    Y = np.zeros_like(X)
    G, gs = pan["groups"], pan["meta_gstride"]
    for p in range(pan["n_panels"]):
        r0, r1 = pan["panel_row_ptr"][p], pan["panel_row_ptr"][p + 1]
        acc = np.zeros((8 * G, X.shape[1]))
        for ch in range(pan["panel_chunk_ptr"][p + 1] - 1, pan["panel_chunk_ptr"][p] - 1, -1):
            h0, h1 = pan["chunk_halo_ptr"][ch], pan["chunk_halo_ptr"][ch + 1]
            staged = X[pan["halo_cols"][h0:h1]]
            k0 = pan["chunk_ks_ptr"][ch]
            m0 = pan["chunk_meta_ptr"][ch]
            gptr = pan["meta"][m0:m0 + gs].astype(np.int64)
            words = pan["meta"][m0 + gs:pan["chunk_meta_ptr"][ch + 1]]
            assert gptr[G] == pan["chunk_ks_ptr"][ch + 1] - k0
            for g in range(G):
                for ks in range(gptr[g], gptr[g + 1]):
                    src = pan["a_src"][32 * (k0 + ks):32 * (k0 + ks + 1)].reshape(8, 4)
                    a = np.where(src >= 0, vals[np.maximum(src, 0)], 0.0)
                    idx = [(int(words[ks]) >> (8 * j)) & 0xff for j in range(4)]
                    assert max(idx) < h1 - h0
                    acc[8 * g:8 * g + 8] += a @ staged[idx]
            if ch == pan["panel_chunk_ptr"][p]:
                assert np.array_equal(pan["halo_cols"][h0:h0 + (r1 - r0)], np.arange(r0, r1))
        Y[r0:r1] = acc[:r1 - r0]
        assert not np.any(acc[r1 - r0:])
    return Y


@pytest.mark.parametrize("name,limits", [("3d_p1", (12, 104, 8, 800)), ("3d_p2", (12, 104, 8, 800)), ("2d_p1", (12, 104, 8, 336)),
                                         ("3d_p1", (12, 52, 10, 800)), ("3d_p2", (12, 48, 16, 800)), ("2d_p2", (12, 52, 10, 400)),
                                         ("3d_p1", (2, 40, 6, 96)), ("2d_p2", (1, 30, 8, 400))])
def test_mma_panels_reproduce_spmm(name, limits):
    """the 8-row-group form (k_spmm_mma) uses every CSR entry exactly once and its dense 8 x 4 blocks times the staged rows
    equal A @ X"""
    import scipy.sparse as sp
    from cases import make_case
    from pygimli_b200.host_setup import build_pattern
    mesh, _, _ = make_case(name)
    rowptr, colidx, _ = build_pattern(mesh)
    groups, hc, nch, rowb = limits
    pan = _capi.build_mma_panels(rowptr, colidx, groups, hc, nch, rowb)
    N = mesh.node_count
    pr = pan["panel_row_ptr"]
    assert pr[0] == 0 and pr[-1] == N and np.all(np.diff(pr) > 0) and np.all(np.diff(pr) <= 8 * groups)
    assert np.all(np.diff(pan["chunk_halo_ptr"]) <= hc) and np.all(np.diff(pan["panel_chunk_ptr"]) <= nch)
    assert np.all(pan["chunk_meta_ptr"] % 4 == 0)
    used = pan["a_src"][:32 * pan["n_ks"]]
    assert np.array_equal(np.sort(used[used >= 0]), np.arange(colidx.size))       # every CSR slot exactly once
    rng = np.random.default_rng(0)
    vals = rng.standard_normal(colidx.size)
    X = rng.standard_normal((N, 5))
    A = sp.csr_matrix((vals, colidx, rowptr), shape=(N, N))
    Y = _replay_mma_spmm(pan, vals, X)
    ref = A @ X
    assert np.max(np.abs(Y - ref)) <= 1e-12 * np.max(np.abs(ref))
