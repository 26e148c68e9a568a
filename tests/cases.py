"""Seeded small test cases shared by the parity tests, smoke() and the golden-vector generator.
Sizes are chosen so that the compiled reference finishes each case in seconds."""
import numpy as np

from pygimli_b200.mesh import (graded_axis, grid_mesh_2d, grid_mesh_3d, create_p2, create_h2, mark_electrode_nodes)
from pygimli_b200.scheme import create_dd, create_slm, create_grid_dd, geometric_factors

CASES = ("2d_p1", "2d_p2", "3d_p1", "3d_p2", "2d_p1_h2", "3d_p1_cellmodel")
EXTRA_CASES = ("3d_crosshole",)      # checked against the reference run live (no committed vectors)


def _model(M, seed=1234):
    rng = np.random.default_rng(seed)
    return 10.0 ** (2.0 + 0.5 * rng.standard_normal(M))


def make_case(name: str):
    """-> (MeshArrays, SchemeArrays with analytic k, model vector)"""
    if name.startswith("2d"):
        ne, sp = 11, 1.0
        xs = graded_axis(0.0, (ne - 1) * sp, sp / 2, 1.4, 60.0)
        ys = -graded_axis(0.0, 4.0, sp / 2, 1.4, 60.0, both=False)
        mesh = grid_mesh_2d(xs, ys, para_box=(-2.0, (ne - 1) * sp + 2.0, -5.0))
        sens = np.zeros((ne, 3))
        sens[:, 0] = np.arange(ne) * sp
        if name == "2d_p1_h2":
            mesh = create_h2(mesh)
        mark_electrode_nodes(mesh, sens)
        if name == "2d_p2":
            mesh = create_p2(mesh)
        scheme = create_dd(sens) if name != "2d_p2" else create_slm(sens)
        scheme.k = geometric_factors(scheme, 2)
    elif name == "3d_crosshole":
        # two short boreholes with buried electrodes (mirror sources with z < 0, SURVEY §8 C4 in miniature)
        h = 1.0
        xs = graded_axis(-2.0, 8.0, h, 1.6, 60.0)
        zs = -graded_axis(0.0, 8.0, h, 1.6, 60.0, both=False)
        mesh = grid_mesh_3d(xs, xs, zs, para_box=(-2.5, 8.5, -2.5, 8.5, -8.5), marker_per="cube")
        sens = np.array([[bx, 3.0, -1.0 - i] for bx in (1.0, 5.0) for i in range(5)], float)
        ids = mark_electrode_nodes(mesh, sens)
        assert np.all(ids >= 0)
        rows = []
        for i in range(4):
            for j in range(4):
                rows.append((i, i + 1, 5 + j, 5 + j + 1))                 # cross-hole dipoles
        rows += [(0, 1, 3, 4), (5, 6, 8, 9), (0, 4, 5, 9)]                   # in-hole and long dipoles
        r = np.asarray(rows, np.int32)
        from pygimli_b200.scheme import SchemeArrays
        scheme = SchemeArrays(sens, r[:, 0], r[:, 1], r[:, 2], r[:, 3])
        scheme.k = geometric_factors(scheme, 3)
    else:
        nx = 5
        sp = 2.0
        h = sp if name == "3d_p2" else sp / 2
        xs = graded_axis(0.0, (nx - 1) * sp, h, 1.6, 50.0)
        zs = -graded_axis(0.0, 4.0, h, 1.6, 50.0, both=False)
        pb = (-2.5, (nx - 1) * sp + 2.5, -2.5, (nx - 1) * sp + 2.5, -4.5)
        mesh = grid_mesh_3d(xs, xs, zs, para_box=pb, marker_per="cube")
        gx, gy = np.meshgrid(np.arange(nx) * sp, np.arange(nx) * sp)
        sens = np.stack([gx.ravel(), gy.ravel(), np.zeros(nx * nx)], 1)
        mark_electrode_nodes(mesh, sens)
        if name == "3d_p2":
            mesh = create_p2(mesh)
        scheme = create_grid_dd(nx, nx, sens)
        scheme.k = geometric_factors(scheme, 3)
    M = int(mesh.cell_marker.max()) + 1
    model = _model(mesh.cell_count if name.endswith("cellmodel") else M)
    return mesh, scheme, model


# Cases sized so that the kernels bench.py times are the ones compared with the reference:
#   3d_p1_wide  72 surface electrodes, complete dipole-dipole: > 64 source columns -> two column tiles of the NC = 2 panel
#               SpMM in all three epilogue roles inside the CUDA graph, a 72 x 72 Gram block
#   2d_p1_wide  65 electrodes, 2.5-D: column tiles that straddle two wavenumber groups (the two_k path of the panel SpMM)
#   3d_p1_192   192 surface electrodes, every 7th row of the complete dipole-dipole scheme: the current-electrode list
#               does not fit one Gram block -> several Jacobian chunks, two register tiles per thread
WIDE_CASES = ("3d_p1_wide", "2d_p1_wide", "3d_p1_192")


def make_wide_case(name: str):
    """-> (MeshArrays, SchemeArrays with analytic k, model, rows): ``rows`` = the data rows whose Jacobian rows the golden
    file holds (the reference computes only those; every row of J depends on the potentials alone)"""
    from pygimli_b200.scheme import create_dd_complete
    if name == "2d_p1_wide":
        ne, sp = 65, 1.0
        xs = graded_axis(0.0, (ne - 1) * sp, sp / 2, 1.4, 300.0)
        ys = -graded_axis(0.0, 16.0, sp / 2, 1.4, 300.0, both=False)
        mesh = grid_mesh_2d(xs, ys, para_box=(-2.0, (ne - 1) * sp + 2.0, -17.0))
        sens = np.zeros((ne, 3))
        sens[:, 0] = np.arange(ne) * sp
        mark_electrode_nodes(mesh, sens)
        scheme = create_dd(sens)
        scheme.k = geometric_factors(scheme, 2)
    else:
        nx, ny = (9, 8) if name == "3d_p1_wide" else (16, 12)
        sp = 1.0
        xs = graded_axis(0.0, (nx - 1) * sp, sp, 1.6, 60.0)
        ysx = graded_axis(0.0, (ny - 1) * sp, sp, 1.6, 60.0)
        zs = -graded_axis(0.0, 4.0, sp, 1.6, 60.0, both=False)
        pb = (-1.5, (nx - 1) * sp + 1.5, -1.5, (ny - 1) * sp + 1.5, -4.5)
        mesh = grid_mesh_3d(xs, ysx, zs, para_box=pb, marker_per="cube")
        gx, gy = np.meshgrid(np.arange(nx) * sp, np.arange(ny) * sp)
        sens = np.stack([gx.ravel(), gy.ravel(), np.zeros(nx * ny)], 1)
        mark_electrode_nodes(mesh, sens)
        scheme = create_dd_complete(sens)
        if name == "3d_p1_192":
            scheme = scheme.subset(np.arange(0, scheme.size, 7))
        scheme.k = geometric_factors(scheme, 3)
        ok = np.isfinite(scheme.k) & (np.abs(scheme.k) < 1e9)
        if not ok.all():
            scheme = scheme.subset(np.nonzero(ok)[0])
    M = int(mesh.cell_marker.max()) + 1
    rows = np.unique(np.linspace(0, scheme.size - 1, 48).astype(int))
    return mesh, scheme, _model(M, seed=4321), rows


def coverage_case(dim: int, seed: int = 77):
    """-> (parameter mesh with one cell per model entry and shuffled markers, dense J, dd, mm, response, model):
    seeded inputs of the coverage tests (coverageDCtrans / createCoverage, bertJacobian.cpp:569-628)"""
    rng = np.random.default_rng(seed + dim)
    if dim == 2:
        mesh = grid_mesh_2d(graded_axis(0.0, 6.0, 1.0, 1.3, 0.0), -graded_axis(0.0, 3.0, 0.5, 1.3, 0.0, both=False))
    else:
        ax = graded_axis(0.0, 3.0, 1.0, 1.3, 0.0)
        mesh = grid_mesh_3d(ax, ax, -graded_axis(0.0, 2.0, 1.0, 1.3, 0.0, both=False))
    M = mesh.cell_count
    mesh.cell_marker = rng.permutation(M).astype(np.int32)
    D = 37
    J = rng.standard_normal((D, M)) * 10.0 ** rng.uniform(-6, 0, size=(D, 1))
    dd = rng.standard_normal(D)
    mm = rng.standard_normal(M) + 3.0
    resp = 10.0 ** rng.uniform(1, 3, D)
    model = 10.0 ** rng.uniform(1, 3, M)
    return mesh, J, dd, mm, resp, model


TOPO_CASES = ("topo_2d", "topo_3d")


def make_topo_case(name: str):
    """flat case with a smooth hill pressed into it: surface faces get different centre heights -> the reference's
    topography branch (numeric primary potentials from a P2 solve, numeric geometric factors).
    -> (MeshArrays, SchemeArrays WITHOUT k, model vector)"""
    mesh, scheme, model = make_case("2d_p1" if name == "topo_2d" else "3d_p1")
    v = mesh.dim - 1                                   # vertical coordinate: y in 2-D meshes, z in 3-D
    x = mesh.pos[:, 0]
    bump = 0.6 * np.exp(-((x - 4.3) / 3.0) ** 2)
    if mesh.dim == 3:
        bump = bump * np.exp(-((mesh.pos[:, 1] - 3.1) / 4.0) ** 2)
    depth = -mesh.pos[:, v]
    mesh.pos[:, v] += bump * np.clip(1.0 - depth / 6.0, 0.0, 1.0)      # fades out 6 m below the surface
    mesh._cache.clear()
    el = np.nonzero(mesh.node_marker == -99)[0]
    sens = scheme.sensors.copy()
    for i in range(sens.shape[0]):                     # electrodes ride on their (moved) surface nodes
        j = el[np.argmin(np.abs(mesh.pos[el, 0] - sens[i, 0]) + (np.abs(mesh.pos[el, 1] - sens[i, 1]) if mesh.dim == 3 else 0.0))]
        sens[i] = mesh.pos[j]
    from pygimli_b200.scheme import SchemeArrays
    return mesh, SchemeArrays(sens, scheme.a, scheme.b, scheme.m, scheme.n, None), model


FEM_CASES = ("2d_p1", "2d_p2", "3d_p1", "3d_p2")


def fem_inputs(mesh, seed: int = 31):
    """seeded per-cell coefficients (stiffness a, mass b) and two probe vectors for the generic FEM matrices"""
    rng = np.random.default_rng(seed)
    a = 10.0 ** rng.uniform(-1, 1, mesh.cell_count)
    b = 10.0 ** rng.uniform(-1, 1, mesh.cell_count)
    X = rng.standard_normal((mesh.node_count, 2))
    return a, b, X


def make_pole_case():
    """2d_p1 mesh with pole-dipole, dipole-pole and pole-pole rows (electrode index -1 = electrode at infinity,
    DataMap::data datamap.cpp:195-215, geometricFactors bertMisc.cpp:131-176) -> (mesh, scheme without k, model)"""
    mesh, scheme, model = make_case("2d_p1")
    ne = scheme.sensors.shape[0]
    rows = []
    for i in range(0, ne - 3):
        rows.append((i, -1, i + 1, i + 2))          # pole-dipole
        rows.append((i, i + 1, i + 3, -1))          # dipole-pole
    for i in range(0, ne - 4, 2):
        rows.append((i, -1, i + 4, -1))             # pole-pole
    r = np.asarray(rows, np.int32)
    from pygimli_b200.scheme import SchemeArrays
    return mesh, SchemeArrays(scheme.sensors, r[:, 0], r[:, 1], r[:, 2], r[:, 3], None), model


COMPLEX_CASES = ("2d_p1", "3d_p1")


def complex_model(model, seed: int = 7):
    """complex resistivities for the complex-resistivity (induced polarisation) cases: the real model with phases of
    -10 ... -20 mrad per model cell"""
    rng = np.random.default_rng(seed)
    return model * np.exp(-0.01j * (1.0 + rng.random(model.size)))


def make_tank_case(ref_node: bool = True, calib_node: bool = True):
    """closed 3-D tank (every boundary face homogeneous Neumann): 8 electrodes on the top face, a
    reference-electrode node (-999) and a calibration node (-1000) as in doc/examples/3_ert/plot_modTank3d.py:50-63
    (dcfemmodelling.cpp:1009-1064).  -> (mesh, scheme without k, model)"""
    from pygimli_b200.scheme import SchemeArrays
    xs = np.linspace(0.0, 1.0, 9)
    zs = -np.linspace(0.0, 0.5, 5)
    mesh = grid_mesh_3d(xs, xs, zs, para_box=(-1.0, 2.0, -1.0, 2.0, -1.0), marker_per="cube")
    mesh.bound_marker[:] = -1                                  # closed box
    sens = np.array([[0.25, 0.25, 0.0], [0.5, 0.25, 0.0], [0.75, 0.25, 0.0], [0.75, 0.5, 0.0],
                     [0.75, 0.75, 0.0], [0.5, 0.75, 0.0], [0.25, 0.75, 0.0], [0.25, 0.5, 0.0]])
    ids = mark_electrode_nodes(mesh, sens)
    assert np.all(ids >= 0)

    def node_at(p):
        return int(np.argmin(np.sum((mesh.pos - np.asarray(p)) ** 2, axis=1)))
    if ref_node:
        mesh.node_marker[node_at([0.5, 0.5, -0.5])] = -999
    if calib_node:
        mesh.node_marker[node_at([0.875, 0.125, -0.25])] = -1000
    rows = np.array([(0, 1, 2, 3), (0, 1, 4, 5), (1, 2, 5, 6), (2, 3, 6, 7), (3, 4, 7, 0), (0, 3, 1, 6), (1, 5, 2, 3), (0, 1, 6, 7),
                     (7, 1, 5, 3), (6, 5, 1, 2)], np.int32)
    scheme = SchemeArrays(sens, rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3])
    M = int(mesh.cell_marker.max()) + 1
    rng = np.random.default_rng(11)
    model = 10.0 ** (1.5 + 0.3 * rng.standard_normal(M))
    return mesh, scheme, model
