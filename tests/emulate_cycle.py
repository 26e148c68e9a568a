"""CPU emulation of the multilevel-preconditioned PCG (test infrastructure, not product code).

    python tests/emulate_cycle.py c3 [--scale 0.6] [--theta 0.25] [--theta 0.0] [--k-index 0]

Builds the workload, lets the compiled reference (oracle/_ref) assemble S(rho) and S(1) in the device's node order,
builds the aggregation hierarchy with the product's host code (pygimli_b200.amg_setup + pgb200_pairwise_aggregate) and
runs the same V(1,1) cycle the GPU runs (damped Jacobi with omega = 1.6 / max(2, Gershgorin), 8 coarsest sweeps,
piecewise-constant transfer) inside a scalar PCG to 1e-12.  The iteration counts track the GPU's closely (c3: 178
emulated, 186 on the B200 with theta = 0.25; 330 / 348 with theta = 0), so aggregation ideas can be screened without
GPU time.  Needs /root/reference only through the prebuilt oracle/_ref."""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref  # noqa: E402
from pygimli_b200 import _capi, amg_setup, host_setup as hs, workloads  # noqa: E402


def system(which: str, scale: float, k_index: int = 0):
    r = workloads.WORKLOADS[which](scale)
    mesh, scheme, kw = r[0], r[1], (r[3] if len(r) > 3 else None)
    mesh2 = hs.renumber_nodes(mesh, hs.node_ordering(mesh))
    R = ref.RefERT(mesh2, scheme, sr=True, solver="pcg")
    if kw is not None:
        R.set_kw(kw[0], kw[1])
    k, _ = R.kw()
    rho = R.mapped_model(workloads.model_for(int(mesh.cell_marker.max()) + 1))
    rp, ci = R.pattern()
    vals, _ = R.assemble(float(k[k_index]), rho, boundary=True)
    vals1, _ = R.assemble(float(k[0]), np.ones(mesh.cell_count), boundary=True)     # hierarchy: rho = 1, smallest k
    N = mesh.node_count
    b = np.zeros(N)
    b[N // 3], b[N // 2] = 1.0, -1.0
    b += 1e-3 * np.random.default_rng(0).standard_normal(N)
    return rp, ci, vals, vals1, sp.csr_matrix((vals, ci, rp), shape=(N, N)), b


def iterations(rp, ci, vals, vals1, A, b, theta: float, sweeps: int = 8, tol: float = 1e-12, maxit: int = 5000):
    lv = amg_setup.build_hierarchy(rp, ci, vals1, lambda a, c, v: _capi.pairwise_aggregate(a, c, v, theta=theta))
    mats, v = [A], vals
    for L in lv:
        v = amg_setup._sum_values(v, L["gal_ptr"], L["gal_idx"])
        mats.append(sp.csr_matrix((v, L["colidx"], L["rowptr"]), shape=(L["n"], L["n"])))
    dws = [(1.6 / max(2.0, (abs(M).sum(1).A1 / M.diagonal()).max())) / M.diagonal() for M in mats]

    def vc(l, r):
        Al, dw = mats[l], dws[l]
        if l == len(lv):
            x = dw * r
            for _ in range(sweeps - 1):
                x = x + dw * (r - Al @ x)
            return x
        L = lv[l]
        x = dw * r
        x = x + vc(l + 1, np.add.reduceat((r - Al @ x)[L["mem_idx"]], L["mem_ptr"][:-1]))[L["agg"]]
        return x + dw * (r - Al @ x)
    x, r = np.zeros_like(b), b.copy()
    z = vc(0, r)
    p, rz, bb = z.copy(), r @ z, np.sqrt(b @ b)
    for it in range(1, maxit + 1):
        Ap = A @ p
        al = rz / (p @ Ap)
        x += al * p
        r -= al * Ap
        if np.sqrt(r @ r) <= tol * bb:
            break
        z = vc(0, r)
        rzn = r @ z
        p = z + (rzn / rz) * p
        rz = rzn
    return it, [M.shape[0] for M in mats], sum(M.nnz for M in mats) / mats[0].nnz


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=sorted(workloads.WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--theta", type=float, action="append")
    ap.add_argument("--k-index", type=int, default=0)
    args = ap.parse_args()
    S = system(args.workload, args.scale, args.k_index)
    print(f"{args.workload} scale {args.scale}: N = {S[4].shape[0]}, nnz = {S[4].nnz}", flush=True)
    for theta in (args.theta or [_capi.AGGREGATION_THETA]):
        t0 = time.time()
        it, sizes, oc = iterations(*S, theta=theta)
        print(f"theta {theta:.2f}: levels {sizes}, operator complexity {oc:.2f}, PCG iterations {it} ({time.time() - t0:.1f} s)", flush=True)


if __name__ == "__main__":
    main()
