"""The pg.Mesh / pg.DataContainerERT adapters of the drop-in boundary (SURVEY.md §8(b) "Input types"), exercised with
stand-ins that expose exactly the accessors of the pygimli binding the adapters call (pgcore is not importable in this image):
Mesh.positions/nodeMarkers/cells/cellMarkers/boundaries/dim/cellCount, Node.id, Cell.nodes, Boundary.nodes/marker;
DataContainerERT.sensors, data[token], haveData.  CPU-only."""
import numpy as np
import pytest

from cases import make_case
from pygimli_b200.ert_modelling import _as_mesh, _as_scheme, CoreB200
from pygimli_b200.host_setup import build_plan


class _Node:
    def __init__(self, i):
        self._i = i

    def id(self):
        return self._i


class _Ent:
    def __init__(self, ids, marker=0):
        self._n = [_Node(int(i)) for i in ids]
        self._m = int(marker)

    def nodes(self):
        return self._n

    def marker(self):
        return self._m


class FakePgMesh:
    def __init__(self, m, extra_unmarked_bounds=3):
        self._m = m
        self._bounds = [_Ent(b, mk) for b, mk in zip(m.bounds, m.bound_marker)]
        # a pg.Mesh also carries its unmarked inner faces (marker 0): the adapter must drop them
        self._bounds += [_Ent(m.bounds[i], 0) for i in range(extra_unmarked_bounds)]

    def dim(self):
        return self._m.dim

    def cellCount(self):
        return self._m.cell_count

    def positions(self):
        return self._m.pos

    def nodeMarkers(self):
        return self._m.node_marker

    def cells(self):
        return [_Ent(c) for c in self._m.cells]

    def cellMarkers(self):
        return self._m.cell_marker

    def boundaries(self):
        return self._bounds


class FakeDataContainerERT:
    def __init__(self, s, with_k=True):
        self._s, self._k = s, with_k

    def sensors(self):
        return self._s.sensors

    def __getitem__(self, t):
        return {"a": self._s.a, "b": self._s.b, "m": self._s.m, "n": self._s.n, "k": self._s.k}[t].astype(float)   # tokens are doubles

    def haveData(self, t):
        return t in "abmn" or (t == "k" and self._k)


@pytest.mark.parametrize("name", ["2d_p1", "3d_p1", "2d_p2"])
def test_pg_mesh_adapter_roundtrip(name):
    mesh, scheme, _ = make_case(name)
    got = _as_mesh(FakePgMesh(mesh))
    assert got.dim == mesh.dim and np.array_equal(got.pos, mesh.pos) and np.array_equal(got.cells, mesh.cells)
    assert np.array_equal(got.node_marker, mesh.node_marker) and np.array_equal(got.cell_marker, mesh.cell_marker)
    assert np.array_equal(got.bounds, mesh.bounds) and np.array_equal(got.bound_marker, mesh.bound_marker)      # marker-0 faces dropped
    # the plan built from the adapted inputs is the plan of the native inputs
    P, Q = build_plan(mesh, scheme), build_plan(got, _as_scheme(FakeDataContainerERT(scheme)))
    for nm in ("rowptr", "colidx", "node_perm", "dir_nodes"):
        assert np.array_equal(getattr(P, nm), getattr(Q, nm)), nm
    assert np.array_equal(P.k, Q.k) and P.M == Q.M and P.nE == Q.nE


def test_data_container_adapter():
    mesh, scheme, _ = make_case("2d_p1")
    s = _as_scheme(FakeDataContainerERT(scheme))
    for t in "abmn":
        assert getattr(s, t).dtype == np.int32 and np.array_equal(getattr(s, t), getattr(scheme, t))
    assert np.array_equal(s.k, scheme.k) and np.array_equal(s.sensors, scheme.sensors)
    assert _as_scheme(FakeDataContainerERT(scheme, with_k=False)).k is None
    with pytest.raises(TypeError):
        _as_scheme(object())
    with pytest.raises(TypeError):
        _as_mesh(object())


def test_core_accepts_pg_like_objects_without_a_gpu():
    """setMesh / setData / kValues go through the adapters and the geometry-only plan; no CUDA call is made"""
    mesh, scheme, _ = make_case("2d_p1")
    core = CoreB200(sr=True)
    core.setMesh(FakePgMesh(mesh))
    core.setData(FakeDataContainerERT(scheme))
    k = core.kValues()
    assert k.size == build_plan(mesh, scheme).k.size and np.all(k > 0)
