"""Pins the oracle's topography branch (numeric primary potentials from a P2 total-field solve, numeric geometric
factors; oracle/ert_oracle.py) to the reference's own outputs in tests/golden/topo_2d.npz (tests/make_golden_topo.py).
The 3-D case takes the oracle a minute (sparse direct P2 solve in scipy), so it is checked on the GPU side only. CPU-only."""
import os

import numpy as np
import pytest

from cases import make_case, make_topo_case
from oracle.ert_oracle import OracleERT, numeric_primary
from pygimli_b200 import _capi
from pygimli_b200.host_setup import build_plan
from pygimli_b200.mesh import create_p2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.fixture(scope="module")
def topo():
    mesh, scheme, model = make_topo_case("topo_2d")
    g = np.load(os.path.join(GOLD, "topo_2d.npz"))
    prim = numeric_primary(mesh, scheme, create_p2(mesh), g["k"], g["w"])
    return mesh, scheme, model, g, prim, OracleERT(mesh, scheme, g["k"], g["w"], primary=prim)


def test_topography_detected(topo):
    mesh, scheme, _, _, _, O = topo
    assert O.topography
    assert build_plan(mesh, scheme).topography
    flat_mesh, flat_scheme, _ = make_case("2d_p1")
    assert not OracleERT(flat_mesh, flat_scheme).topography
    assert not build_plan(flat_mesh, flat_scheme).topography


def test_numeric_primary_matches_reference(topo):
    _, _, _, g, prim, _ = topo
    assert _rel(prim[[0, prim.shape[0] - 1]], g["prim_rows"]) < 1e-10


def test_numeric_geometric_factors(topo):
    _, _, _, g, _, O = topo
    assert _rel(O.numeric_geometric_factors(), g["kfac"]) < 1e-10


def test_response_and_jacobian(topo):
    _, scheme, model, g, _, O = topo
    scheme.k = g["kfac"]
    rhoa = O.response(model)
    assert np.max(np.abs(rhoa - g["rhoa"]) / np.abs(g["rhoa"])) < 1e-8
    assert _rel(O.pots[[0, O.pots.shape[0] - 1]], g["pot_rows"]) < 1e-9
    assert _rel(O.jacobian(model), g["J"]) < 1e-9


def test_forward_without_primary_is_refused():
    mesh, scheme, model = make_topo_case("topo_2d")
    with pytest.raises(RuntimeError):
        OracleERT(mesh, scheme).forward(model)


def test_pure_neumann_3d_plan():
    """no mixed/Dirichlet face at all: topography branch, the reference's node 0 becomes the calibration node and the last
    electrode the current reference (dcfemmodelling.cpp:1040-1075)"""
    mesh, scheme, _ = make_case("3d_p1")
    mesh.bound_marker[:] = -1
    P = build_plan(mesh, scheme)
    assert P.topography and P.neumann_domain and P.ref_last == 1 and P.ref_node == -1
    assert list(P.dir_nodes) == [int(P.node_inv[0])]