"""Generate tests/golden/complex_*.npz by RUNNING THE REFERENCE's complex-resistivity path (oracle/_ref:
DCMultiElectrodeModelling with setComplex(true), core/src/bert/dcfemmodelling.cpp:1103-1118, 1199-1208, 1410-1461,
1755-1925) on the seeded cases of tests/cases.py.  Complex solves: scipy SuperLU + 2 refinement steps through the
reference's setSolver seam (SolverWrapper::setMatrix(CSparseMatrix) / solve(CVector), solverWrapper.h:34-40).

    python tests/make_golden_complex.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from cases import COMPLEX_CASES, complex_model, make_case  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    for name in COMPLEX_CASES:
        mesh, scheme, model = make_case(name)
        mc = complex_model(model)
        R = ref.RefERTComplex(mesh, scheme)
        rhoa = R.response(mc)
        U = R.solutions()
        J = R.create_jacobian(mc)
        np.savez_compressed(os.path.join(HERE, "golden", "complex_" + name + ".npz"), rhoa=rhoa, J=J, sol_rows=U[[0, U.shape[0] - 1]])
        print(name, rhoa[:2], J.shape)


if __name__ == "__main__":
    main()
