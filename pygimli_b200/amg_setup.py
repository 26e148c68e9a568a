"""Geometry-only set-up of the multilevel (aggregation) preconditioner of the block-PCG.

The reference solves S x = b with a sparse direct factorisation (CHOLMOD,
core/src/cholmodWrapper.cpp:359-418); the B200 path iterates, and plain Jacobi-PCG needs
thousands of iterations on graded ERT meshes.  The preconditioner is one V(1,1) cycle of an
unsmoothed-aggregation multigrid:

  * aggregates come from two passes of greedy pairwise matching along the strongest negative
    coupling of the rho = 1 matrix (geometry only -> built once per mesh, here on the host);
  * prolongation is piecewise constant, so every coarse matrix is a plain sum of finer-level
    entries: coarse_vals[slot] = sum(fine_vals[gal_idx[gal_ptr[slot]:gal_ptr[slot+1]]]) --
    a gather the GPU redoes for every new resistivity model (and every wavenumber);
  * smoother: damped Jacobi; coarsest level: a fixed number of Jacobi sweeps.

Everything here is index bookkeeping; the numerics run in libpgb200_ert.so.
"""
from __future__ import annotations

import numpy as np


def _coarsen(rowptr, colidx, agg, nc):
    """pattern of P^T A P for piecewise-constant P (agg: fine node -> coarse node) and the gather lists"""
    n = rowptr.size - 1
    rowof = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
    key = agg[rowof].astype(np.int64) * nc + agg[colidx]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    first = np.ones(ks.size, bool)
    first[1:] = ks[1:] != ks[:-1]
    ukey = ks[first]
    gal_ptr = np.concatenate([np.nonzero(first)[0], [ks.size]]).astype(np.int32)
    gal_idx = order.astype(np.int32)
    crow = (ukey // nc).astype(np.int64)
    ccol = (ukey % nc).astype(np.int32)
    crowptr = np.zeros(nc + 1, np.int32)
    np.cumsum(np.bincount(crow, minlength=nc), out=crowptr[1:])
    return crowptr, ccol, gal_ptr, gal_idx


def _sum_values(vals, gal_ptr, gal_idx):
    # sequential sums in gather order (np.bincount accumulates in index order), the order csrc/plan_builder.cpp and the
    # k_galerkin kernel use; np.add.reduceat would switch to pairwise blocks for segments of 8 and more entries, and the
    # last-bit differences flip ties of the strength-based matching
    seg = np.repeat(np.arange(gal_ptr.size - 1), np.diff(gal_ptr))
    return np.bincount(seg, weights=vals[gal_idx], minlength=gal_ptr.size - 1)


def _diag_pos(rowptr, colidx):
    n = rowptr.size - 1
    rowof = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
    d = np.nonzero(colidx == rowof)[0]
    assert d.size == n, "every row needs a diagonal entry"
    return d.astype(np.int32)


def build_hierarchy(rowptr, colidx, vals, aggregate_fn, passes: int = 2, min_size: int = 256, max_levels: int = 12,
                    panel_ptr=None):
    """-> list of coarse levels (finest first).  Each level dict describes the transfer from the
    next finer level: n, nnz, rowptr, colidx, diag_pos, gal_ptr, gal_idx, agg (finer node -> node),
    mem_ptr/mem_idx (members of every aggregate).  ``panel_ptr`` confines the first level's aggregates to the SpMM
    row panels (numbered panel by panel, ``panel_agg_ptr``); measured to cost 2.4x more PCG iterations on graded
    meshes because the strongest couplings cross panel borders, so the product does not use it."""
    levels = []
    rp, ci, v = np.asarray(rowptr, np.int32), np.asarray(colidx, np.int32), np.asarray(vals, np.float64)
    while len(levels) < max_levels:
        n = rp.size - 1
        if n <= min_size:
            break
        agg = np.arange(n, dtype=np.int32)
        rp_p, ci_p, v_p, nc = rp, ci, v, n
        first = not levels and panel_ptr is not None
        group = np.repeat(np.arange(len(panel_ptr) - 1, dtype=np.int32), np.diff(panel_ptr)) if first else None
        for _ in range(passes):
            a, na = aggregate_fn(rp_p, ci_p, v_p, group) if first else aggregate_fn(rp_p, ci_p, v_p)
            if first:
                g2 = np.zeros(na, np.int32)
                g2[a] = group
                group = g2
            crp, cci, gp, gi = _coarsen(rp_p, ci_p, a, na)
            v_p = _sum_values(v_p, gp, gi)
            rp_p, ci_p = crp, cci
            agg = a[agg]
            nc = na
        if nc > 0.7 * n:
            break
        panel_agg_ptr = None
        if first:
            # number the aggregates panel by panel (fine rows are consecutive inside a panel)
            order_a = np.argsort(group, kind="stable")
            newid = np.empty(nc, np.int32)
            newid[order_a] = np.arange(nc, dtype=np.int32)
            agg = newid[agg]
            panel_agg_ptr = np.concatenate([[0], np.cumsum(np.bincount(group, minlength=len(panel_ptr) - 1))]).astype(np.int32)
        crp, cci, gp, gi = _coarsen(rp, ci, agg, nc)
        order = np.argsort(agg, kind="stable").astype(np.int32)
        mem_ptr = np.zeros(nc + 1, np.int32)
        np.cumsum(np.bincount(agg, minlength=nc), out=mem_ptr[1:])
        levels.append(dict(n=int(nc), nnz=int(cci.size), rowptr=crp, colidx=cci, diag_pos=_diag_pos(crp, cci),
                           gal_ptr=gp, gal_idx=gi, agg=agg.astype(np.int32), mem_ptr=mem_ptr, mem_idx=order,
                           panel_agg_ptr=panel_agg_ptr))
        v = _sum_values(v, gp, gi)
        rp, ci = crp, cci
    return levels
