"""ctypes front-end of ``oracle/_ref/libgimli_ref.so`` -- TEST INFRASTRUCTURE ONLY.

The shared object is the UNMODIFIED reference C++ (built by ``oracle/Makefile`` from
/root/reference/core/src) plus ``oracle/ref_driver.cpp``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU arm may import this module.

The reference's linear solver (CHOLMOD, absent from this image) is replaced through the
reference's own ``setSolver`` seam by scipy's SuperLU (a sparse *direct* solver) followed by
two steps of iterative refinement -- independent of the CUDA PCG it is used to check.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libgimli_ref.so")

_SETM = C.CFUNCTYPE(None, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double))
_SOLVE = C.CFUNCTYPE(None, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double))

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_create.restype = C.c_void_p
        _lib.ref_refine.restype = C.c_void_p
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _d(a):
    return _p(a, C.c_double)


def _i(a):
    return _p(a, C.c_int)


class DirectSolver:
    """scipy SuperLU + iterative refinement; stands in for CHOLMOD (cholmodWrapper.cpp:359-418)."""

    def __init__(self):
        self.lu = None
        self.A = None
        self.n_factor = 0
        self.n_solve = 0

    def set_matrix(self, n, nnz, rowptr, colidx, vals):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        rp = np.ctypeslib.as_array(rowptr, shape=(n + 1,)).copy()
        ci = np.ctypeslib.as_array(colidx, shape=(nnz,)).copy()
        v = np.ctypeslib.as_array(vals, shape=(nnz,)).copy()
        self.A = sp.csr_matrix((v, ci, rp), shape=(n, n))
        self.lu = spla.splu(self.A.tocsc(), permc_spec="MMD_AT_PLUS_A",
                            diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
        self.n_factor += 1

    def solve(self, n, rhs, sol):
        b = np.ctypeslib.as_array(rhs, shape=(n,))
        x = self.lu.solve(b)
        for _ in range(2):
            x = x + self.lu.solve(b - self.A @ x)
        np.ctypeslib.as_array(sol, shape=(n,))[:] = x
        self.n_solve += 1


class ComplexDirectSolver:
    """scipy SuperLU on the complex-symmetric system + 2 refinement steps (stands in for CHOLMOD/UMFPACK complex,
    cholmodWrapper.cpp:167-222)"""

    def __init__(self):
        self.lu = None
        self.n_factor = 0
        self.n_solve = 0

    def set_matrix(self, n, nnz, rowptr, colidx, vals):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        rp = np.ctypeslib.as_array(rowptr, shape=(n + 1,)).copy()
        ci = np.ctypeslib.as_array(colidx, shape=(nnz,)).copy()
        v = np.ctypeslib.as_array(vals, shape=(2 * nnz,)).copy().view(np.complex128)
        self.A = sp.csr_matrix((v, ci, rp), shape=(n, n))
        self.lu = spla.splu(self.A.tocsc())
        self.n_factor += 1

    def solve(self, n, rhs, sol):
        b = np.ctypeslib.as_array(rhs, shape=(2 * n,)).view(np.complex128)
        x = self.lu.solve(b)
        for _ in range(2):
            x = x + self.lu.solve(b - self.A @ x)
        np.ctypeslib.as_array(sol, shape=(2 * n,))[:] = np.ascontiguousarray(x).view(np.float64)
        self.n_solve += 1


class RefERTComplex:
    """The reference's complex-resistivity path: ``DCMultiElectrodeModelling`` with ``setComplex(True)``
    (total field; model = [re(rho); im(rho)], response = [re(rhoa); im(rhoa)], complex Jacobian)."""

    def __init__(self, mesh, scheme):
        L = lib()
        L.ref_create_complex.restype = C.c_void_p
        self.mesh, self.scheme = mesh, scheme
        pos = np.ascontiguousarray(mesh.pos, np.float64)
        nm = np.ascontiguousarray(mesh.node_marker, np.int32)
        cells = np.ascontiguousarray(mesh.cells, np.int32)
        cm = np.ascontiguousarray(mesh.cell_marker, np.int32)
        bounds = np.ascontiguousarray(mesh.bounds, np.int32)
        bm = np.ascontiguousarray(mesh.bound_marker, np.int32)
        sens = np.ascontiguousarray(scheme.sensors, np.float64)
        abmn = scheme.abmn()
        self.h = C.c_void_p(L.ref_create_complex(
            C.c_int(mesh.dim), C.c_int(pos.shape[0]), _d(pos), _i(nm),
            C.c_int(cells.shape[0]), C.c_int(cells.shape[1]), _i(cells), _i(cm),
            C.c_int(bounds.shape[0]), C.c_int(bounds.shape[1] if bounds.ndim == 2 else 0), _i(bounds), _i(bm),
            C.c_int(sens.shape[0]), _d(sens), C.c_int(abmn.shape[0]), _i(abmn)))
        self.N, self.D, self.nE = pos.shape[0], abmn.shape[0], sens.shape[0]
        self.solver = ComplexDirectSolver()
        self._cb1 = _SETM(self.solver.set_matrix)
        self._cb2 = _SOLVE(self.solver.solve)
        L.ref_set_complex_callbacks(self.h, self._cb1, self._cb2)
        L.ref_set_threads(self.h, C.c_int(1))
        if scheme.k is not None:
            k = np.ascontiguousarray(scheme.k, np.float64)
            L.ref_set_k(self.h, _d(k))

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def kw(self):
        n = lib().ref_n_k(self.h)
        k, w = np.zeros(n), np.zeros(n)
        lib().ref_get_kw(self.h, _d(k), _d(w))
        return k, w

    def response(self, model_c):
        """model_c: complex cell/model resistivities -> complex apparent resistivities"""
        mc = np.asarray(model_c, np.complex128)
        m = np.ascontiguousarray(np.concatenate([mc.real, mc.imag]))
        out = np.zeros(2 * self.D)
        lib().ref_response_complex(self.h, C.c_int(m.size), _d(m), _d(out))
        return out[: self.D] + 1j * out[self.D:]

    def solutions(self):
        """k-summed potentials, rows [re(e = 0..nE-1); im(e = 0..nE-1)] -> complex (nE, N)"""
        r = lib().ref_solution_rows(self.h)
        out = np.zeros((r, self.N))
        lib().ref_get_solutions(self.h, _d(out))
        return out[: self.nE] + 1j * out[self.nE:]

    def create_jacobian(self, model_c):
        mc = np.asarray(model_c, np.complex128)
        m = np.ascontiguousarray(np.concatenate([mc.real, mc.imag]))
        rc = np.zeros(2, np.int32)
        lib().ref_create_jacobian_complex(self.h, C.c_int(m.size), _d(m), _i(rc), None)
        J = np.zeros((int(rc[0]), 2 * int(rc[1])))
        lib().ref_create_jacobian_complex(self.h, C.c_int(m.size), _d(m), _i(rc), _d(J))
        return J.view(np.complex128)


class RefERT:
    """The reference's ``DCSRMultiElectrodeModelling`` (sr=True) or ``DCMultiElectrodeModelling``
    on flat arrays.  ``mesh``/``scheme`` are duck-typed (pygimli_b200.MeshArrays / SchemeArrays)."""

    def __init__(self, mesh, scheme, sr=True, solver="direct", verbose=False):
        L = lib()
        self.mesh, self.scheme = mesh, scheme
        pos = np.ascontiguousarray(mesh.pos, np.float64)
        nm = np.ascontiguousarray(mesh.node_marker, np.int32)
        cells = np.ascontiguousarray(mesh.cells, np.int32)
        cm = np.ascontiguousarray(mesh.cell_marker, np.int32)
        bounds = np.ascontiguousarray(mesh.bounds, np.int32)
        bm = np.ascontiguousarray(mesh.bound_marker, np.int32)
        sens = np.ascontiguousarray(scheme.sensors, np.float64)
        abmn = scheme.abmn()
        self.h = C.c_void_p(L.ref_create(
            C.c_int(mesh.dim), C.c_int(pos.shape[0]), _d(pos), _i(nm),
            C.c_int(cells.shape[0]), C.c_int(cells.shape[1]), _i(cells), _i(cm),
            C.c_int(bounds.shape[0]), C.c_int(bounds.shape[1] if bounds.ndim == 2 else 0), _i(bounds), _i(bm),
            C.c_int(sens.shape[0]), _d(sens), C.c_int(abmn.shape[0]), _i(abmn),
            C.c_int(1 if sr else 0), C.c_int(1 if verbose else 0)))
        self.N = pos.shape[0]
        self.Cn = cells.shape[0]
        self.D = abmn.shape[0]
        self.nE = sens.shape[0]
        self.solver = None
        if solver == "direct":
            self.solver = DirectSolver()
            self._cb1 = _SETM(self.solver.set_matrix)
            self._cb2 = _SOLVE(self.solver.solve)
            L.ref_set_solver_callbacks(self.h, self._cb1, self._cb2)
        if scheme.k is not None:
            self.set_k(scheme.k)

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # configuration -------------------------------------------------------
    def set_threads(self, n):
        lib().ref_set_threads(self.h, C.c_int(n))

    def kw(self):
        n = lib().ref_n_k(self.h)
        k, w = np.zeros(n), np.zeros(n)
        lib().ref_get_kw(self.h, _d(k), _d(w))
        return k, w

    def set_kw(self, k, w):
        k = np.ascontiguousarray(k, np.float64)
        w = np.ascontiguousarray(w, np.float64)
        lib().ref_set_kw(self.h, C.c_int(k.size), _d(k), _d(w))

    def fill_matrix(self, kind, coef):
        """CSR values of fillStiffnessMatrix (kind 0) / fillMassMatrix (kind 1), pattern as ``pattern()``"""
        coef = np.ascontiguousarray(coef, np.float64)
        n = lib().ref_fill_matrix(self.h, C.c_int(kind), _d(coef), None)
        out = np.zeros(n)
        lib().ref_fill_matrix(self.h, C.c_int(kind), _d(coef), _d(out))
        return out

    def set_primary_from(self, p2_handle: "RefERT"):
        """numeric primary potentials from a total-field run on the P2 mesh (see ref_driver.cpp ref_set_primary_from)"""
        return int(lib().ref_set_primary_from(self.h, p2_handle.h))

    def topography(self):
        return bool(lib().ref_topography(self.h))

    def electrode_nodes(self):
        n = lib().ref_n_electrodes(self.h)
        out = np.zeros(n, np.int32)
        lib().ref_electrode_nodes(self.h, _i(out))
        return out

    def geometric_factors(self):
        out = np.zeros(self.D)
        lib().ref_geometric_factors(self.h, _d(out))
        return out

    def set_k(self, k):
        k = np.ascontiguousarray(k, np.float64)
        lib().ref_set_k(self.h, _d(k))

    # the path -------------------------------------------------------------
    def response(self, model):
        m = np.ascontiguousarray(model, np.float64)
        out = np.zeros(self.D)
        lib().ref_response(self.h, C.c_int(m.size), _d(m), _d(out))
        return out

    def mapped_model(self, model):
        m = np.ascontiguousarray(model, np.float64)
        out = np.zeros(self.Cn)
        lib().ref_mapped_model(self.h, C.c_int(m.size), _d(m), _d(out))
        return out

    def clear_potentials(self):
        lib().ref_clear_potentials(self.h)

    def create_jacobian(self, model):
        m = np.ascontiguousarray(model, np.float64)
        rc = np.zeros(2, np.int32)
        lib().ref_create_jacobian(self.h, C.c_int(m.size), _d(m), _i(rc))
        J = np.zeros((int(rc[0]), int(rc[1])))
        lib().ref_get_jacobian(self.h, _d(J))
        return J

    def sensitivity_only(self, pots, n_threads=1, want=True):
        pots = np.ascontiguousarray(pots, np.float64)
        rc = np.zeros(2, np.int32)
        M = int(self.mesh.cell_marker.max()) + 1
        J = np.zeros((self.D, M)) if want else None
        lib().ref_sensitivity_only(self.h, C.c_int(n_threads), C.c_int(pots.shape[0]), _d(pots),
                                   _d(J) if want else None, _i(rc))
        return J

    def subpotentials(self):
        r = lib().ref_subpot_rows(self.h)
        out = np.zeros((r, self.N))
        if r:
            lib().ref_get_subpotentials(self.h, _d(out))
        return out

    def solutions(self):
        out = np.zeros((self.nE, self.N))
        lib().ref_get_solutions(self.h, _d(out))
        return out

    def primary(self):
        r = lib().ref_get_primary(self.h, None)
        out = np.zeros((r, self.N))
        if r:
            lib().ref_get_primary(self.h, _d(out))
        return out

    def pattern(self):
        nnz = lib().ref_pattern(self.h, None, None)
        rp = np.zeros(self.N + 1, np.int32)
        ci = np.zeros(nnz, np.int32)
        lib().ref_pattern(self.h, _i(rp), _i(ci))
        return rp, ci

    def assemble(self, k, rho_cells, boundary=True, want=True):
        rho = np.ascontiguousarray(rho_cells, np.float64)
        nnz = lib().ref_pattern(self.h, None, None) if want else 0
        vals = np.zeros(nnz) if want else None
        sec = np.zeros(2)
        lib().ref_assemble(self.h, C.c_double(k), _d(rho), C.c_int(1 if boundary else 0),
                           _d(vals) if want else None, _d(sec))
        return vals, sec

    def element_matrices(self, cell):
        n = self.mesh.cells.shape[1]
        K, M = np.zeros((n, n)), np.zeros((n, n))
        lib().ref_element_matrices(self.h, C.c_int(int(cell)), _d(K), _d(M))
        return K, M

    def time_forward_jacobian(self, model, n_threads):
        m = np.ascontiguousarray(model, np.float64)
        out = np.zeros(2)
        lib().ref_time_forward_jacobian(self.h, C.c_int(m.size), _d(m), C.c_int(n_threads), _d(out))
        return out

    def set_pcg_tol(self, tol):
        lib().ref_set_pcg_tol(self.h, C.c_double(tol))

    def time_partial_solve(self, model, n_src, detail=False):
        """seconds the reference spends in SolverWrapper::solve for the first n_src sources (all k)"""
        m = np.ascontiguousarray(model, np.float64)
        out = np.zeros(4)
        lib().ref_time_partial_solve(self.h, C.c_int(m.size), _d(m), C.c_int(int(n_src)), _d(out))
        return out if detail else float(out[1])

    def partial_solve_pots(self, model, n_src):
        """(times[4], pots[n_src * nK, N]): the reference's calculateK for the first n_src sources, every wavenumber;
        row i + k * n_src holds the total potential of source i at wavenumber k"""
        m = np.ascontiguousarray(model, np.float64)
        out = np.zeros(4)
        nK = lib().ref_n_k(self.h)
        sol = np.zeros((int(n_src) * nK, self.N))
        lib().ref_partial_solve_pots(self.h, C.c_int(m.size), _d(m), C.c_int(int(n_src)), _d(out), _d(sol))
        return out, sol

    def solver_stats(self):
        out = np.zeros(4)
        lib().ref_solver_stats(self.h, _d(out))
        return dict(n_set_matrix=int(out[0]), t_set_matrix=out[1], n_solve=int(out[2]), t_solve=out[3])


def bessel(x):
    x = np.ascontiguousarray(x, np.float64)
    k0, k1 = np.zeros_like(x), np.zeros_like(x)
    lib().ref_bessel(C.c_int(x.size), _d(x), _d(k0), _d(k1))
    return k0, k1


def kwave_list(rmin, rmax, nleg, nlag):
    k, w = np.zeros(nleg + nlag), np.zeros(nleg + nlag)
    lib().ref_kwave_list(C.c_double(rmin), C.c_double(rmax), C.c_int(nleg), C.c_int(nlag), _d(k), _d(w))
    return k, w


def refine(mesh, kind):
    """reference createH2 (kind=1) / createP2 (kind=2) -> dict of flat arrays"""
    L = lib()
    pos = np.ascontiguousarray(mesh.pos, np.float64)
    nm = np.ascontiguousarray(mesh.node_marker, np.int32)
    cells = np.ascontiguousarray(mesh.cells, np.int32)
    cm = np.ascontiguousarray(mesh.cell_marker, np.int32)
    bounds = np.ascontiguousarray(mesh.bounds, np.int32)
    bm = np.ascontiguousarray(mesh.bound_marker, np.int32)
    h = C.c_void_p(L.ref_refine(C.c_int(mesh.dim), C.c_int(pos.shape[0]), _d(pos), _i(nm),
                                C.c_int(cells.shape[0]), C.c_int(cells.shape[1]), _i(cells), _i(cm),
                                C.c_int(bounds.shape[0]), C.c_int(bounds.shape[1]), _i(bounds), _i(bm), C.c_int(kind)))
    sz = np.zeros(5, np.int32)
    L.ref_mesh_sizes(h, _i(sz))
    out = dict(pos=np.zeros((sz[0], 3)), node_marker=np.zeros(sz[0], np.int32),
               cells=np.zeros((sz[1], sz[2]), np.int32), cell_marker=np.zeros(sz[1], np.int32),
               bounds=np.zeros((sz[3], sz[4]), np.int32), bound_marker=np.zeros(sz[3], np.int32))
    L.ref_mesh_export(h, _d(out["pos"]), _i(out["node_marker"]), _i(out["cells"]), _i(out["cell_marker"]),
                      _i(out["bounds"]), _i(out["bound_marker"]))
    return out


def coverage_trans(J, dd, mm):
    """the reference's coverageDCtrans (core/src/bert/bertJacobian.cpp:569) on a dense matrix"""
    J = np.ascontiguousarray(J, np.float64)
    dd = np.ascontiguousarray(dd, np.float64)
    mm = np.ascontiguousarray(mm, np.float64)
    out = np.zeros(J.shape[1])
    lib().ref_coverage_trans(C.c_int(J.shape[0]), C.c_int(J.shape[1]), _d(J), _d(dd), _d(mm), _d(out))
    return out


def create_coverage(J, mesh, response=None, model=None):
    """the reference's createCoverage (core/src/bert/bertJacobian.cpp:600-628); ``mesh`` is the parameter mesh"""
    J = np.ascontiguousarray(J, np.float64)
    pos = np.ascontiguousarray(mesh.pos, np.float64)
    cells = np.ascontiguousarray(mesh.cells, np.int32)
    cm = np.ascontiguousarray(mesh.cell_marker, np.int32)
    out = np.zeros(cells.shape[0])
    r = None if response is None else np.ascontiguousarray(response, np.float64)
    m = None if model is None else np.ascontiguousarray(model, np.float64)
    lib().ref_create_coverage(C.c_int(J.shape[0]), C.c_int(J.shape[1]), _d(J), C.c_int(mesh.dim), C.c_int(pos.shape[0]), _d(pos),
                              C.c_int(cells.shape[0]), C.c_int(cells.shape[1]), _i(cells), _i(cm),
                              _d(r) if r is not None else None, _d(m) if m is not None else None, _d(out))
    return out
