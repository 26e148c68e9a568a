"""Generate tests/golden/tank_*.npz by RUNNING THE REFERENCE (oracle/_ref) on the closed-tank cases of tests/cases.py
(make_tank_case): pure-Neumann 3-D domain, reference-electrode node (-999), calibration node (-1000) or node 0
(core/src/bert/dcfemmodelling.cpp:141-161, 1009-1064, 1517-1523), total-field operator (DCMultiElectrodeModelling),
numeric geometric factors (:1539-1556).

    python tests/make_golden_tank.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from cases import make_tank_case  # noqa: E402
from oracle import ref  # noqa: E402
from pygimli_b200.scheme import SchemeArrays  # noqa: E402

CASES = {"tank_ref_cal": (True, True), "tank_ref": (True, False), "tank_last": (False, True)}


def main():
    for name, (rn, cn) in CASES.items():
        mesh, scheme, model = make_tank_case(rn, cn)
        sch1 = SchemeArrays(scheme.sensors, scheme.a, scheme.b, scheme.m, scheme.n, np.ones(scheme.size))
        R = ref.RefERT(mesh, sch1, sr=False, solver="direct")
        R.set_threads(1)
        assert R.topography()
        k = R.geometric_factors()
        R.set_k(k)
        rhoa = R.response(model)
        pots = R.subpotentials()
        out = dict(kfac=k, rhoa=rhoa, pots=pots)
        if rn:
            out["J"] = R.create_jacobian(model)       # without a reference node the reference's createJacobian throws
        np.savez_compressed(os.path.join(HERE, "golden", name + ".npz"), **out)
        print(name, {k_: v.shape for k_, v in out.items()})


if __name__ == "__main__":
    main()
