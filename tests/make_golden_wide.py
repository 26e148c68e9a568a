"""Generate tests/golden/wide_*.npz: the reference itself (oracle/_ref) on the cases that are big enough to run the
kernels bench.py times (tests/cases.py WIDE_CASES).  The reference solves all sources (potentials, rhoa of the full
scheme); its sensitivity loop runs on a subset of the rows only (cost is linear in the rows, every row depends on the
potentials alone).

    python tests/make_golden_wide.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from cases import WIDE_CASES, make_wide_case  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    for name in WIDE_CASES:
        mesh, scheme, model, rows = make_wide_case(name)
        R = ref.RefERT(mesh, scheme, sr=True, solver="direct")
        R.set_threads(1)
        k, w = R.kw()
        rhoa = R.response(model)
        pots = R.subpotentials()
        R.close()
        Rs = ref.RefERT(mesh, scheme.subset(rows), sr=True, solver="direct")
        Rs.set_threads(1)
        J = Rs.create_jacobian(model)
        Rs.close()
        nS = pots.shape[0]
        pick = sorted(set([0, 1, nS // 3, nS // 2, nS - 2, nS - 1]))
        path = os.path.join(HERE, "golden", "wide_" + name + ".npz")
        np.savez_compressed(path, k=k, w=w, rhoa=rhoa, pots_rows=np.asarray(pick), pots=pots[pick], rows=rows, J=J,
                            model=model, kfac=scheme.k)
        print(name, "N", mesh.node_count, "C", mesh.cell_count, "nE", scheme.sensor_count, "D", scheme.size, "nK", k.size,
              "M", model.size, "->", os.path.getsize(path) // 1024, "KiB", flush=True)


if __name__ == "__main__":
    main()
