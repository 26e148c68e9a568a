"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (oracle/_ref, compiled from
/root/reference) on the seeded cases of tests/cases.py.  Run here (the container with the
reference sources); the vectors travel to the GPU box as committed fixtures.

    python tests/make_golden.py

Settings: setThreadCount(1) (the threaded sensitivity loop has a benign write race,
bertJacobian.cpp:233); linear solves by scipy SuperLU + 2 refinement steps through the
reference's setSolver seam (CHOLMOD is not installed).  Potentials are stored for a few
sources only to keep the fixtures small.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from cases import CASES, make_case  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    os.makedirs(os.path.join(HERE, "golden"), exist_ok=True)
    for name in CASES:
        mesh, scheme, model = make_case(name)
        R = ref.RefERT(mesh, scheme, sr=True, solver="direct")
        R.set_threads(1)
        k, w = R.kw()
        rp, ci = R.pattern()
        rho = R.mapped_model(model)
        vals0, _ = R.assemble(float(k[0]), rho, boundary=True)
        valsL, _ = R.assemble(float(k[-1]), rho, boundary=True)
        rhoa = R.response(model)
        pots = R.subpotentials()
        J = R.create_jacobian(model)
        R.clear_potentials()
        hom = np.full(model.size, 100.0)
        Jh = R.create_jacobian(hom)          # analytic branch (dcfemmodelling.cpp:1272-1301)
        nS = pots.shape[0]
        pick = sorted(set([0, nS // 2, nS - 1]))
        out = dict(k=k, w=w, rowptr=rp, colidx=ci, rho=rho, vals_k0=vals0, vals_klast=valsL, rhoa=rhoa,
                   pots_rows=np.asarray(pick), pots=pots[pick], J=J, J_hom=Jh, model=model,
                   el_nodes=R.electrode_nodes(), kfac=scheme.k)
        path = os.path.join(HERE, "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "N", mesh.node_count, "C", mesh.cell_count, "D", scheme.size, "nK", k.size,
              "->", os.path.getsize(path) // 1024, "KiB")
        R.close()


if __name__ == "__main__":
    main()
