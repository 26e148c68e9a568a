"""Flat-array measuring schemes (the reference's ``DataContainerERT``).

The reference stores the tokens ``a b m n k`` as float vectors and the sensor list as
3-vectors (core/src/bert/bertDataContainer.cpp:47-62); index -1 is an unused electrode.
The generators restate the index patterns of pygimli/physics/ert/ertScheme.py
(``dd`` :452-523, ``slm`` :723-749) so that synthetic schemes have the reference's sizes
(41 electrodes dd -> 741 rows, 96 electrodes slm -> 2209 rows).
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np


@dataclass
class SchemeArrays:
    sensors: np.ndarray   # (nE, 3) float64
    a: np.ndarray         # (D,) int32
    b: np.ndarray
    m: np.ndarray
    n: np.ndarray
    k: np.ndarray | None = None   # (D,) float64 geometric factors

    def __post_init__(self):
        self.sensors = np.ascontiguousarray(self.sensors, dtype=np.float64).reshape(-1, 3)
        for t in "abmn":
            setattr(self, t, np.ascontiguousarray(getattr(self, t), dtype=np.int32))
        if self.k is not None:
            self.k = np.ascontiguousarray(self.k, dtype=np.float64)
        ne = self.sensors.shape[0]
        for t in "abmn":
            v = getattr(self, t)
            if v.size and (v.max() >= ne or v.min() < -1):
                raise IndexError(f"electrode index out of range in token {t}")

    @property
    def size(self) -> int:
        return int(self.a.size)

    @property
    def sensor_count(self) -> int:
        return int(self.sensors.shape[0])

    def abmn(self) -> np.ndarray:
        return np.ascontiguousarray(np.stack([self.a, self.b, self.m, self.n], 1), dtype=np.int32)

    def subset(self, idx) -> "SchemeArrays":
        idx = np.asarray(idx)
        return SchemeArrays(self.sensors, self.a[idx], self.b[idx], self.m[idx], self.n[idx],
                            None if self.k is None else self.k[idx])


def _from_rows(sensors, rows) -> SchemeArrays:
    r = np.asarray(rows, dtype=np.int32).reshape(-1, 4)
    return SchemeArrays(sensors, r[:, 0], r[:, 1], r[:, 2], r[:, 3])


def dipole_dipole_rows(ne: int, offset: int = 0, stride: int = 1):
    rows = []
    for sep in range(1, ne):
        for i in range(ne - 1 - sep):
            a, b = i, i + 1
            m = b + sep
            n = m + 1
            if n < ne:
                rows.append((offset + a * stride, offset + b * stride, offset + m * stride, offset + n * stride))
    return rows


def create_dd(sensors) -> SchemeArrays:
    """open dipole-dipole line scheme, dipole length 1 (ertScheme.py:500-521)"""
    sensors = np.asarray(sensors, float).reshape(-1, 3)
    return _from_rows(sensors, dipole_dipole_rows(sensors.shape[0]))


def create_slm(sensors) -> SchemeArrays:
    """Wenner-Schlumberger C--P-P--C (ertScheme.py:737-746)"""
    sensors = np.asarray(sensors, float).reshape(-1, 3)
    ne = sensors.shape[0]
    rows = []
    for sep in range(1, ne - 1):
        for i in range(ne - 2 - sep):
            a = i
            m = a + sep
            n = m + 1
            b = n + sep
            if b < ne:
                rows.append((a, b, m, n))
    return _from_rows(sensors, rows)


def create_dd_complete(sensors) -> SchemeArrays:
    """closed/complete dipole-dipole: every ordered pair of disjoint neighbour dipoles
    (ertScheme.py:484-498 with complete=True)."""
    sensors = np.asarray(sensors, float).reshape(-1, 3)
    ne = sensors.shape[0]
    rows = []
    for i in range(ne):
        a, b = i, (i + 1) % ne
        for j in range(ne):
            m, n = j, (j + 1) % ne
            if a != m and a != n and b != m and b != n:
                rows.append((a, b, m, n))
    return _from_rows(sensors, rows)


def create_grid_dd(nx: int, ny: int, sensors) -> SchemeArrays:
    """3-D surface grid: inline dipole-dipole along every x-line and every y-line
    (10x10 grid -> 20 lines * 28 rows = 560 rows)."""
    rows = []
    for j in range(ny):
        rows += dipole_dipole_rows(nx, offset=j * nx, stride=1)
    for i in range(nx):
        rows += dipole_dipole_rows(ny, offset=i, stride=nx)
    return _from_rows(np.asarray(sensors, float).reshape(-1, 3), rows)


def geometric_factors(scheme: SchemeArrays, dim: int = 3) -> np.ndarray:
    """Analytic flat-earth geometric factors, k = 1/(uAM - uBM - uAN + uBN) with
    u = (1/r + 1/r')/(4 pi) and the mirror source at z -> -z
    (core/src/bert/bertMisc.cpp:131-176, :186-214).  In 2-D a non-zero y is moved to z first."""
    s = scheme.sensors.copy()
    if dim == 2:
        sw = s[:, 1] != 0.0
        s[sw, 2] = s[sw, 1]
        s[sw, 1] = 0.0

    def u(i, j):
        out = np.zeros(i.shape, float)
        ok = (i > -1) & (j > -1)
        p, q = s[j[ok]], s[i[ok]]          # potential at p for source q
        r = np.sqrt(np.sum((p - q) ** 2, 1))
        qm = q.copy()
        qm[:, 2] = -qm[:, 2]
        rm = np.sqrt(np.sum((p - qm) ** 2, 1))
        val = np.where(r < 1e-12, 1.0, (1.0 / np.where(r < 1e-12, 1.0, r) + 1.0 / rm) / (4.0 * np.pi))
        out[ok] = val
        return out

    a, b, m, n = scheme.a, scheme.b, scheme.m, scheme.n
    with np.errstate(divide="ignore"):
        return 1.0 / (u(a, m) - u(b, m) - u(a, n) + u(b, n))


def electrode_matrix_data(pm: np.ndarray, scheme: SchemeArrays) -> np.ndarray:
    """DataMap::data (core/src/bert/datamap.cpp:195-215): four-point voltages from the electrode-potential matrix
    pm[source electrode, pick-up electrode]; electrode index -1 contributes nothing"""
    def P(a, m):
        return np.where((a >= 0) & (m >= 0), pm[np.maximum(a, 0), np.maximum(m, 0)], 0.0)
    return (P(scheme.a, scheme.m) - P(scheme.a, scheme.n)) - (P(scheme.b, scheme.m) - P(scheme.b, scheme.n))
