"""Synthetic workloads named in BASELINE.json `configs`: graded tensor-grid
meshes (no Triangle/TetGen in this image), reference-style schemes, seeded resistivity models.
Used by bench.py and by the full-size property tests; nothing here is on the numeric path."""
from __future__ import annotations

import numpy as np

from .mesh import graded_axis, grid_mesh_2d, grid_mesh_3d, create_p2, mark_electrode_nodes
from .host_setup import kwave_from_range
from .scheme import create_dd, create_slm, create_dd_complete, create_grid_dd, geometric_factors, dipole_dipole_rows


def model_for(M: int, seed: int = 1234) -> np.ndarray:
    """rho = 10^(2 + 0.5 g), g ~ N(0,1) per model cell"""
    rng = np.random.default_rng(seed)
    return 10.0 ** (2.0 + 0.5 * rng.standard_normal(M))


def c1_2d_dd(scale: float = 1.0):
    """configs[0]: 2.5D dipole-dipole, 41-electrode line, ~10k-cell para mesh (P1 triangles)"""
    ne, sp = 41, 1.0
    h = sp / (4.0 * scale)
    xs = graded_axis(-2.0, (ne - 1) * sp + 2.0, h, 1.3, 400.0)
    ys = -graded_axis(0.0, (ne - 1) * sp / 3.0, h, 1.3, 400.0, both=False)
    mesh = grid_mesh_2d(xs, ys, para_box=(-2.0, (ne - 1) * sp + 2.0, -(ne - 1) * sp / 3.0))
    sens = np.zeros((ne, 3))
    sens[:, 0] = np.arange(ne) * sp
    mark_electrode_nodes(mesh, sens)
    scheme = create_dd(sens)
    scheme.k = geometric_factors(scheme, 2)
    return mesh, scheme, "2.5D dd, 41 electrodes, P1 triangles"


def c2_2d_slm_p2(scale: float = 1.0):
    """configs[1]: 2.5D Wenner-Schlumberger, 96 electrodes, ~200k P2 triangles, 11 wavenumbers"""
    ne, sp = 96, 1.0
    h = sp / (4.0 * scale)
    xs = graded_axis(-3.0, (ne - 1) * sp + 3.0, h, 1.25, 1000.0)
    ys = -graded_axis(0.0, 40.0, h, 1.25, 1000.0, both=False)
    mesh = grid_mesh_2d(xs, ys, para_box=(-3.0, (ne - 1) * sp + 3.0, -40.0))
    sens = np.zeros((ne, 3))
    sens[:, 0] = np.arange(ne) * sp
    mark_electrode_nodes(mesh, sens)
    mesh = create_p2(mesh)
    scheme = create_slm(sens)
    scheme.k = geometric_factors(scheme, 2)
    # 11 wavenumbers set explicitly (setkValues/setWeights): 7 Legendre + 4 Laguerre nodes
    kw = kwave_from_range(sp / 2.0, (ne - 1) * sp * 2.0, 7, 4)
    return mesh, scheme, "2.5D slm, 96 electrodes, P2 triangles, 11 wavenumbers", kw


def c3_3d_grid(scale: float = 1.0, complete: bool = True, marker_per: str = "cell"):
    """configs[2]: 3D surface ERT, 10x10 electrode grid (2 m), ~1M tetrahedra, dipole-dipole"""
    nx, sp = 10, 2.0
    h = 0.5 / scale
    lo, hi = -3.0, (nx - 1) * sp + 3.0
    xs = graded_axis(lo, hi, h, 1.5, 250.0)
    zs = -graded_axis(0.0, 9.0, h, 1.5, 250.0, both=False)
    mesh = grid_mesh_3d(xs, xs, zs, para_box=(lo, hi, lo, hi, -9.0), marker_per=marker_per)
    gx, gy = np.meshgrid(np.arange(nx) * sp, np.arange(nx) * sp)
    sens = np.stack([gx.ravel(), gy.ravel(), np.zeros(nx * nx)], 1)
    mark_electrode_nodes(mesh, sens)
    scheme = create_dd_complete(sens) if complete else create_grid_dd(nx, nx, sens)
    scheme.k = geometric_factors(scheme, 3)
    return mesh, scheme, "3D surface, 10x10 electrodes, Tet4, " + ("complete dd" if complete else "inline dd")


def c4_3d_crosshole(scale: float = 1.0):
    """configs[3]: 3D crosshole ERT, 8 boreholes x 24 electrodes on a circle (R = 10 m, 1 m spacing from z = -2),
    ~4M tetrahedra; dipole-dipole within every borehole and between neighbouring boreholes"""
    nb, ne_b, R = 8, 24, 10.0
    h = 0.5 / scale
    ang = 2.0 * np.pi * np.arange(nb) / nb
    # snap the borehole positions to the grid so that electrodes sit on nodes
    bx = np.round(R * np.cos(ang) / h) * h
    by = np.round(R * np.sin(ang) / h) * h
    lo, hi = -R - 6.0, R + 6.0                       # fine region sized for ~4M tetrahedra (BASELINE.json configs[3])
    xs = graded_axis(lo, hi, h, 1.5, 300.0)
    zs = -graded_axis(0.0, 33.0, h, 1.5, 300.0, both=False)
    mesh = grid_mesh_3d(xs, xs, zs, para_box=(lo, hi, lo, hi, -33.0), marker_per="cube")
    sens = np.array([[bx[b], by[b], -2.0 - i] for b in range(nb) for i in range(ne_b)])
    ids = mark_electrode_nodes(mesh, sens)
    assert np.all(ids >= 0), "crosshole electrodes must coincide with grid nodes"
    rows = []
    for b in range(nb):
        rows += dipole_dipole_rows(ne_b, offset=b * ne_b)                      # in-hole
        nb2 = (b + 1) % nb                                                     # cross-hole with the neighbour
        for i in range(ne_b - 1):
            for j in range(0, ne_b - 1, 3):
                rows.append((b * ne_b + i, b * ne_b + i + 1, nb2 * ne_b + j, nb2 * ne_b + j + 1))
    r = np.asarray(rows, np.int32)
    from .scheme import SchemeArrays
    scheme = SchemeArrays(sens, r[:, 0], r[:, 1], r[:, 2], r[:, 3])
    scheme.k = geometric_factors(scheme, 3)
    return mesh, scheme, "3D crosshole, 8 boreholes x 24 electrodes, Tet4"


def c5_3d_timelapse(scale: float = 1.0):
    """configs[4]: the mesh/scheme of one frame of the 3D time-lapse loop: 100 surface electrodes, ~2M tetrahedra
    (the Gauss-Newton loop itself is `steps` repetitions of response + createJacobian in bench.py)"""
    return c3_3d_grid(scale=scale * 1.26, complete=True, marker_per="cube")[:2] + ("3D surface time-lapse frame, 10x10 electrodes, ~2M Tet4",)


WORKLOADS = {"c1": c1_2d_dd, "c2": c2_2d_slm_p2, "c3": c3_3d_grid, "c4": c4_3d_crosshole, "c5": c5_3d_timelapse}
