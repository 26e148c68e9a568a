"""Generate tests/golden/fem.npz by RUNNING THE REFERENCE's SparseMatrix::fillStiffnessMatrix / fillMassMatrix
(core/src/sparsematrix.h:1034-1065, compiled into oracle/_ref) on the seeded cases of tests/cases.py with seeded per-cell
coefficients.  To keep the fixture small, each matrix is stored through its products with two seeded vectors and its
diagonal (every entry enters a product with an independent random weight).

    python tests/make_golden_fem.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from cases import make_case, fem_inputs, FEM_CASES  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    out = {}
    for name in FEM_CASES:
        mesh, scheme, _ = make_case(name)
        a, b, X = fem_inputs(mesh)
        R = ref.RefERT(mesh, scheme, sr=True, solver="direct")
        rp, ci = R.pattern()
        for tag, kind, coef in (("K", 0, a), ("M", 1, b)):
            S = sp.csr_matrix((R.fill_matrix(kind, coef), ci, rp), shape=(mesh.node_count, mesh.node_count))
            out[f"{name}_{tag}x"] = S @ X
            out[f"{name}_{tag}diag"] = S.diagonal()
    np.savez_compressed(os.path.join(HERE, "golden", "fem.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
