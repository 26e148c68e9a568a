"""Generate tests/golden/topo_*.npz by RUNNING THE REFERENCE (oracle/_ref) on the topography cases of tests/cases.py:
numeric primary potentials from the reference's own P2 total-field run (Mesh::createP2, DCMultiElectrodeModelling,
interpolate), handed to DCSRMultiElectrodeModelling::setPrimaryPotential -- the three steps of checkPrimpotentials_
(core/src/bert/dcfemmodelling.cpp:2009-2056) driven from outside because its temporary fop cannot take the injected
solver (no CHOLMOD in this image); numeric geometric factors by calcGeometricFactor (:1539-1556).

    python tests/make_golden_topo.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from cases import TOPO_CASES, make_topo_case  # noqa: E402
from oracle import ref  # noqa: E402
from pygimli_b200.mesh import MeshArrays  # noqa: E402
from pygimli_b200.scheme import SchemeArrays  # noqa: E402


def run_reference(mesh, scheme, model):
    p2 = ref.refine(mesh, 2)
    p2m = MeshArrays(mesh.dim, p2["pos"], p2["node_marker"], p2["cells"], p2["cell_marker"], p2["bounds"], p2["bound_marker"])
    sch1 = SchemeArrays(scheme.sensors, scheme.a, scheme.b, scheme.m, scheme.n, np.ones(scheme.size))
    Rp = ref.RefERT(p2m, sch1, sr=False, solver="direct")
    Rp.set_threads(1)
    assert Rp.topography()
    Rp.response(np.ones(int(mesh.cell_marker.max()) + 1))
    R = ref.RefERT(mesh, sch1, sr=True, solver="direct")
    R.set_threads(1)
    assert R.topography()
    nrows = R.set_primary_from(Rp)
    kv, w = R.kw()
    k = R.geometric_factors()
    R.set_k(k)
    rhoa = R.response(model)
    pots = R.subpotentials()
    J = R.create_jacobian(model)
    R.clear_potentials()
    hom = R.response(np.full(model.size, 100.0))
    prim = np.zeros((nrows, mesh.node_count))
    ref.lib().ref_get_primary(R.h, prim.ctypes.data_as(C.POINTER(C.c_double)))
    return dict(k=kv, w=w, kfac=k, rhoa=rhoa, J=J, rhoa_hom=hom, prim_rows=prim[[0, nrows - 1]], pot_rows=pots[[0, nrows - 1]])


def main():
    for name in TOPO_CASES:
        mesh, scheme, model = make_topo_case(name)
        out = run_reference(mesh, scheme, model)
        np.savez_compressed(os.path.join(HERE, "golden", name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
