"""Generate tests/golden/pole_2d.npz by RUNNING THE REFERENCE (oracle/_ref) on the pole-dipole / pole-pole scheme of
tests/cases.py::make_pole_case (electrode index -1).      python tests/make_golden_pole.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from cases import make_pole_case  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    mesh, scheme, model = make_pole_case()
    R = ref.RefERT(mesh, scheme, sr=True, solver="direct")
    R.set_threads(1)
    k = R.geometric_factors()                       # analytic (flat earth), bertMisc.cpp:131-176
    R.set_k(k)
    rhoa = R.response(model)
    J = R.create_jacobian(model)
    np.savez_compressed(os.path.join(HERE, "golden", "pole_2d.npz"), kfac=k, rhoa=rhoa, J=J)
    print(k[:6], rhoa[:6], J.shape)


if __name__ == "__main__":
    main()
