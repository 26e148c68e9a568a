"""Generate tests/golden/coverage.npz by RUNNING THE REFERENCE's coverageDCtrans / createCoverage
(core/src/bert/bertJacobian.cpp:569-628, compiled into oracle/_ref) on a seeded dense matrix and a small
parameter mesh (one cell per model entry, shuffled markers).  Run here; the vectors travel as fixtures.

    python tests/make_golden_coverage.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from cases import coverage_case  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    out = {}
    for dim in (2, 3):
        mesh, J, dd, mm, resp, model = coverage_case(dim)
        out[f"cov_trans_{dim}d"] = ref.coverage_trans(J, dd, mm)
        out[f"coverage_{dim}d"] = ref.create_coverage(J, mesh, resp, model)
        out[f"coverage_unit_{dim}d"] = ref.create_coverage(J, mesh)
    np.savez_compressed(os.path.join(HERE, "golden", "coverage.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
