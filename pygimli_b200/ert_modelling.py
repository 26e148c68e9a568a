"""Host-side mirror of the reference's ERT forward operator for the B200 path.

``ERTModellingB200`` has the call surface of ``pygimli.physics.ert.ERTModelling``
(pygimli/physics/ert/ertModelling.py:73-246) for the forward + Jacobian path:

    fop = ERTModellingB200(sr=True)
    fop.setMesh(mesh)          # MeshArrays or pg.Mesh (when pygimli is importable)
    fop.setData(scheme)        # SchemeArrays or pg.DataContainerERT      (fop.data = scheme works too)
    rhoa = fop.response(model)
    fop.createJacobian(model)
    J = fop.jacobian()         # rows() cols() mult(x) transMult(y)  -- J stays in HBM

``CoreB200`` is the replacement for ``pg.core.DCSRMultiElectrodeModelling`` behind
``ERTModelling._core`` (same method names: setMesh, setData, response, createJacobian, jacobian,
setkValues, setWeights, kValues, weights, calcGeometricFactor, solution, mapERTModel,
setThreadCount, setVerbose).  Everything numeric runs in libpgb200_ert.so (CUDA, sm_100a);
there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi
from .amg_setup import build_hierarchy
from .host_setup import build_plan, TOLERANCE
from .mesh import MeshArrays, from_pg_mesh, create_p2
from .scheme import SchemeArrays, geometric_factors, electrode_matrix_data


DEFAULT_PCG_TOL = 1e-12          # 3-D (one system per source)
DEFAULT_PCG_TOL_25D = 2e-13      # 2.5-D: the systems of the smallest wavenumbers are close to singular; rows of J whose four
                                 # potential terms nearly cancel need this to stay within 1e-8 of the reference (measured)


def _as_mesh(mesh) -> MeshArrays:
    if isinstance(mesh, MeshArrays):
        return mesh
    if hasattr(mesh, "cellCount") and hasattr(mesh, "positions"):
        return from_pg_mesh(mesh)
    raise TypeError("mesh must be a pygimli_b200.MeshArrays or a pg.Mesh")


def _as_scheme(data) -> SchemeArrays:
    if isinstance(data, SchemeArrays):
        return data
    if hasattr(data, "sensorPositions") or hasattr(data, "sensors"):
        sens = np.asarray(data.sensors() if hasattr(data, "sensors") else data.sensorPositions(), float)
        tok = {t: np.asarray(data[t]).astype(np.int32) for t in "abmn"}
        k = np.asarray(data["k"], float) if data.haveData("k") else None
        return SchemeArrays(sens, tok["a"], tok["b"], tok["m"], tok["n"], k)
    raise TypeError("data must be a pygimli_b200.SchemeArrays or a pg.DataContainerERT")


class JacobianB200:
    """The Jacobian as the inversion sees it (``ModellingBase::jacobian()``,
    core/src/modellingbase.h:118-121): rows(), cols(), mult(x), transMult(y).
    The matrix lives in HBM (column-major); ``numpy()`` copies it out row-major."""

    def __init__(self, core: "CoreB200"):
        self._core = core

    def rows(self) -> int:
        return self._core._jac_shape()[0]

    def cols(self) -> int:
        return self._core._jac_shape()[1]

    @property
    def shape(self):
        return self._core._jac_shape()[:2]

    def mult(self, x):
        x = np.ascontiguousarray(x, np.float64)
        if x.size != self.cols():
            raise ValueError("mult: vector length must equal cols()")
        y = np.zeros(self.rows())
        _capi.check(_capi.lib().pgb200_ert_jacobian_mult(self._core._h, x.ctypes.data, y.ctypes.data))
        return y

    def transMult(self, y):
        y = np.ascontiguousarray(y, np.float64)
        if y.size != self.rows():
            raise ValueError("transMult: vector length must equal rows()")
        x = np.zeros(self.cols())
        _capi.check(_capi.lib().pgb200_ert_jacobian_tmult(self._core._h, y.ctypes.data, x.ctypes.data))
        return x

    # ---- weighted products and coverage: J stays in HBM -------------------
    def _vec(self, v, n, what):
        if v is None:
            return None
        v = np.ascontiguousarray(v, np.float64)
        if v.size != n:
            raise ValueError(f"{what}: vector length {v.size} does not fit ({n})")
        return v

    def mult_lr(self, x, left=None, right=None):
        """left .* (J (right .* x)) -- MultLeftRightMatrix.mult (pygimli/frameworks/inversion.py:705-708)"""
        x = self._vec(x, self.cols(), "mult")
        left, right = self._vec(left, self.rows(), "left"), self._vec(right, self.cols(), "right")
        y = np.zeros(self.rows())
        _capi.check(_capi.lib().pgb200_ert_jacobian_mult_lr(
            self._core._h, left.ctypes.data if left is not None else None, right.ctypes.data if right is not None else None,
            x.ctypes.data, y.ctypes.data))
        return y

    def transMult_lr(self, y, left=None, right=None):
        """right .* (J^T (left .* y)) -- MultLeftRightMatrix.transMult"""
        y = self._vec(y, self.rows(), "transMult")
        left, right = self._vec(left, self.rows(), "left"), self._vec(right, self.cols(), "right")
        x = np.zeros(self.cols())
        _capi.check(_capi.lib().pgb200_ert_jacobian_tmult_lr(
            self._core._h, left.ctypes.data if left is not None else None, right.ctypes.data if right is not None else None,
            y.ctypes.data, x.ctypes.data))
        return x

    def coverageDCtrans(self, dd, mm=None):
        """cov[j] = sum_i |J_ij dd_i| / |mm_j|  (core/src/bert/bertJacobian.cpp:569-598); mm=None -> undivided sums"""
        dd = self._vec(dd, self.rows(), "dd")
        mm = self._vec(mm, self.cols(), "mm")
        cov = np.zeros(self.cols())
        _capi.check(_capi.lib().pgb200_ert_coverage_trans(self._core._h, dd.ctypes.data, mm.ctypes.data if mm is not None else None,
                                                          cov.ctypes.data))
        return cov

    def numpy(self) -> np.ndarray:
        r, c = self.shape
        out = np.zeros((r, c))
        _capi.check(_capi.lib().pgb200_ert_jacobian_copy(self._core._h, out.ctypes.data))
        return out

    __array__ = lambda self, dtype=None, copy=None: self.numpy()  # noqa: E731

    def torch(self):
        """zero-copy torch view of logical shape (rows, cols) onto the column-major HBM buffer"""
        import torch
        ptr, rows, cols, ld = self._core._jac_info()

        class _Iface:
            __cuda_array_interface__ = dict(shape=(cols, ld), typestr="<f8", data=(ptr, False), version=3, strides=None)
        t = torch.as_tensor(_Iface(), device=f"cuda:{self._core.device}")
        return t[:, :rows].t()


class MultLeftRightMatrixB200:
    """``pg.matrix.MultLeftRightMatrix(J, left, right)`` over the HBM-resident Jacobian: the error-/transform-weighted
    Jacobian the Gauss-Newton inversion hands to its LSQR/CG solver (pygimli/frameworks/inversion.py:705-708, 776-779)."""

    def __init__(self, A: JacobianB200, left, right):
        if A.cols() != len(right):
            raise Exception("Matrix columns do not fit right vector length!")
        if A.rows() != len(left):
            raise Exception("Matrix rows do not fit left vector length!")
        self.A = A
        self.l = np.ascontiguousarray(left, np.float64)
        self.r = np.ascontiguousarray(right, np.float64)

    def rows(self):
        return self.A.rows()

    def cols(self):
        return self.A.cols()

    def mult(self, x):
        return self.A.mult_lr(x, self.l, self.r)

    def transMult(self, y):
        return self.A.transMult_lr(y, self.l, self.r)


def coverageDCtrans(S: JacobianB200, dd, mm):
    """drop-in for ``pg.core.coverageDCtrans(S, dd, mm)`` (core/src/bert/bertJacobian.cpp:569)"""
    return S.coverageDCtrans(dd, mm)


def createCoverage(S: JacobianB200, mesh, response=None, model=None):
    """drop-in for ``pg.core.createCoverage(S, mesh[, response, model])`` (core/src/bert/bertJacobian.cpp:600-628):
    coverageDCtrans(S, 1/response, 1/model) looked up per cell of ``mesh`` (the parameter domain, markers 0..M-1) and
    divided by the cell sizes.  The reference's branch for cellCount != len(model) cannot succeed (it sizes the
    per-marker accumulator by cells and then requires every entry to be positive), so that case raises here."""
    mesh = _as_mesh(mesh)
    response = np.ones(S.rows()) if response is None else np.asarray(response, float)
    model = np.ones(S.cols()) if model is None else np.asarray(model, float)
    cov_model = S.coverageDCtrans(1.0 / response, 1.0 / model)
    marker = np.asarray(mesh.cell_marker)
    if marker.min() < 0 or marker.max() >= cov_model.size:
        raise IndexError("createCoverage: cell markers of the mesh must index the model vector")
    if mesh.cell_count != model.size:
        raise RuntimeError(f"Coverage fails:{mesh.cell_count} {model.size}")
    return cov_model[marker] / mesh.cell_sizes()


def managerCoverage(S, para_mesh, response, model):
    """``ERTManager.coverage()`` (pygimli/physics/ert/ertManager.py:328-340): coverage under the logarithmic
    transformation, log10(coverageDCtrans(J, 1/response, 1/model) / parameter sizes), looked up per cell of the parameter
    mesh (markers 0..M-1, several cells may share a marker).  ``S`` is the HBM-resident Jacobian (or anything with a
    ``coverageDCtrans(dd, mm)`` method); only M doubles leave the GPU."""
    para_mesh = _as_mesh(para_mesh)
    response, model = np.asarray(response, float), np.asarray(model, float)
    cov_trans = S.coverageDCtrans(1.0 / response, 1.0 / model)
    marker = np.asarray(para_mesh.cell_marker)
    if marker.min() < 0 or marker.max() >= model.size:
        raise IndexError("coverage: cell markers of the parameter mesh must index the model vector")
    param_sizes = np.bincount(marker, weights=para_mesh.cell_sizes(), minlength=model.size)
    with np.errstate(divide="ignore"):
        return np.log10(cov_trans / param_sizes)[marker]


class ComplexJacobianB200:
    """Complex sensitivity matrix of the complex-resistivity path (the reference's ``CMatrix`` jacobian,
    dcfemmodelling.cpp:1446-1461): ``numpy()`` is the D x M complex matrix, ``squeezed(conj)`` the real 2D x 2M block matrix
    ``pg.utils.squeezeComplex`` / ``toRealMatrix`` hands to the inversion (pygimli/utils/complex.py:79-111)."""

    def __init__(self, J):
        self._J = J

    def rows(self):
        return self._J.shape[0]

    def cols(self):
        return self._J.shape[1]

    def numpy(self):
        return self._J

    def squeezed(self, conj=False):
        re, im = self._J.real, self._J.imag
        return np.block([[re, im], [-im, re]]) if conj else np.block([[re, -im], [im, re]])


class CoreB200:
    """Replacement for ``pg.core.DCSRMultiElectrodeModelling`` (sr=True) /
    ``DCMultiElectrodeModelling`` (sr=False) on one B200."""

    def __init__(self, sr: bool = True, verbose: bool = False, device: int = 0, preconditioner: str = "multilevel"):
        if preconditioner not in ("multilevel", "jacobi"):
            raise ValueError("preconditioner must be 'multilevel' or 'jacobi'")
        self.preconditioner = preconditioner
        self.plan_builder = os.environ.get("PGB200_PLAN", "native")
        self.hierarchy = None
        self.sr = bool(sr)
        self.verbose = bool(verbose)
        self.device = int(device)
        self._mesh = None
        self._scheme = None
        self._plan = None
        self._h = None
        self._keep = None
        self._k = None
        self._w = None
        # stated relative residual tolerance of the block-PCG, ||r|| <= tol * ||b|| per source column (PGB200_TOL overrides
        # the default for tolerance studies)
        self._tol, self._maxit, self._check = None, 50000, 25       # None: the default of the problem's dimension, see _ensure_handle
        self._stream = None
        self._warm = False
        self._shard = None
        self._prim_pm = None
        self._placeholder_k = False
        self._J = JacobianB200(self)
        self._complex = False       # complex resistivity (setComplex): an inner total-field core on the doubled electrode list
        self._cx = None
        self._Jc = None

    # ---- reference-style setters -----------------------------------------------------
    def setVerbose(self, v):
        self.verbose = bool(v)

    def setThreadCount(self, n):   # CPU threads of the reference's sensitivity loop; nothing to do on the GPU
        pass

    def setComplex(self, c=True):
        """complex resistivity (induced polarisation), DCMultiElectrodeModelling::setComplex: models and responses are
        [real part | imaginary part] vectors, the Jacobian is complex (ertModelling.py:204-238, dcfemmodelling.cpp:1103-1118,
        1446-1461).  Total-field fop only: the reference's singularity-removal class refuses complex models (:2156)."""
        if c and self.sr:
            raise _capi.PGB200Error("complex resistivity needs the total-field operator (sr=False); the reference's "
                                    "DCSRMultiElectrodeModelling::calculateK throws for complex models")
        self._complex = bool(c)
        self._drop_complex()

    def complex(self):
        return self._complex

    def _drop_complex(self):
        if self._cx is not None:
            self._cx.close()
        self._cx = None
        self._Jc = None

    def _ensure_complex(self):
        """inner core: same mesh, every electrode listed twice (source column i: real part, i + nE: imaginary part), scheme =
        the four real blocks of the complex sensitivity, wavenumbers of the ORIGINAL electrode list (include/pgb200_ert.h)"""
        if self._cx is None:
            P = self._ensure_plan()                       # wavenumbers and weights of the original layout
            sch = self._scheme
            nE = sch.sensors.shape[0]

            def off(v):
                return np.where(v >= 0, v + nE, v)
            a, b, m, n = sch.a, sch.b, sch.m, sch.n
            sch4 = SchemeArrays(np.vstack([sch.sensors, sch.sensors]),
                                np.concatenate([a, off(a), a, off(a)]), np.concatenate([b, off(b), b, off(b)]),
                                np.concatenate([m, off(m), off(m), m]), np.concatenate([n, off(n), off(n), n]),
                                np.ones(4 * sch.size))
            cx = CoreB200(sr=False, verbose=self.verbose, device=self.device)
            cx.setMesh(self._mesh)
            cx.setData(sch4)
            cx.setkValues(P.k)
            cx.setWeights(P.w)
            cx.setSolverTolerance(self._tol, self._maxit, self._check)
            h = cx._ensure_handle()
            _capi.check(_capi.lib().pgb200_ert_set_complex(h, 1))
            self._cx = cx
        return self._cx

    def _complex_model(self, model):
        m = np.ascontiguousarray(model, np.float64).ravel()
        if m.size % 2:
            raise _capi.PGB200Error("complex model: expected [real part | imaginary part]")
        return m, m.size // 2

    def setMesh(self, mesh, ignoreRegionManager=True):
        self._mesh = _as_mesh(mesh)
        self._invalidate()

    def setData(self, data):
        self._scheme = _as_scheme(data)
        self._invalidate()

    def setkValues(self, k):
        self._k = np.asarray(k, float).copy()
        self._invalidate()

    def setWeights(self, w):
        self._w = np.asarray(w, float).copy()
        self._invalidate()

    def kValues(self):
        self._ensure_plan()
        return self._plan.k.copy()

    def weights(self):
        self._ensure_plan()
        return self._plan.w.copy()

    def setSolverTolerance(self, rel_tol=None, max_iter=50000, check_every=25):
        """block-PCG controls (replaces the reference's direct CHOLMOD solve; stated tolerance)"""
        self._tol, self._maxit, self._check = (None if rel_tol is None else float(rel_tol)), int(max_iter), int(check_every)
        if self._h:
            if self._tol is None:
                self._tol = DEFAULT_PCG_TOL if self._plan.dim == 3 else DEFAULT_PCG_TOL_25D
            _capi.check(_capi.lib().pgb200_ert_set_solver(self._h, self._tol, self._maxit, self._check))

    def setWarmStart(self, on=True):
        """start every block-PCG solve from the potentials of the previous one (Gauss-Newton / time-lapse loops)"""
        self._warm = bool(on)
        if self._h:
            _capi.check(_capi.lib().pgb200_ert_set_warm_start(self._h, 1 if self._warm else 0))

    def setStream(self, stream_ptr):
        """CUDA stream handle (int, e.g. torch.cuda.current_stream().cuda_stream)"""
        self._stream = int(stream_ptr) if stream_ptr else None
        if self._h:
            _capi.check(_capi.lib().pgb200_ert_set_stream(self._h, C.c_void_p(self._stream or 0)))

    def setShard(self, src_begin, src_end, row_begin, row_end):
        self._shard = (int(src_begin), int(src_end), int(row_begin), int(row_end))
        if self._h:
            _capi.check(_capi.lib().pgb200_ert_set_shard(self._h, *self._shard))

    def calcGeometricFactor(self, data=None, nModel=0):
        """geometric factors (dcfemmodelling.cpp:1527-1556): analytic for flat earth; with topography numeric,
        1 / (u(rho = 1) + TOLERANCE) from the electrode potentials of a rho = 1 solve"""
        sch = self._scheme if data is None else _as_scheme(data)
        topo = False
        if self._mesh is not None and self._scheme is not None:
            topo = self._ensure_plan().topography
        if not topo:
            return geometric_factors(sch, self._mesh.dim if self._mesh is not None else 3)
        self._ensure_handle()
        if self.sr:
            pm = self._prim_pm                       # SR with rho = 1: secondary field is zero, u = primary potentials
        else:
            # total field: u(rho = 1) from a solve on this mesh (dcfemmodelling.cpp:1539-1556).  The k-factors are what is
            # being computed, so the forward call must not ask for them (response() would raise without them, :1096)
            P = self._plan
            placeholder, self._placeholder_k = self._placeholder_k, False
            try:
                self.response(np.ones(nModel if nModel > 0 else P.M))
            finally:
                self._placeholder_k = placeholder
            pm = self.get("pm").reshape(P.nE, P.nE)
            self.clearPotentials()
        return 1.0 / (electrode_matrix_data(pm, sch) + TOLERANCE)

    # ---- life cycle -------------------------------------------------------------------
    def _invalidate(self):
        self._drop_complex()
        if self._h:
            _capi.lib().pgb200_ert_destroy(self._h)
        self._h = None
        if isinstance(self._plan, _capi.NativePlan):
            self._plan.free()
        self._plan = None
        self._keep = None
        self._prim_pm = None
        self._placeholder_k = False

    def close(self):
        self._invalidate()

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

    def _ensure_plan(self):
        if self._plan is None:
            if self._mesh is None:
                raise RuntimeError("Found no mesh, so cannot calculate a response.")
            if self._scheme is None:
                raise RuntimeError("no response without data container")
            # the plan is geometry only; missing k-factors are resolved at the first response()/createJacobian().
            # Built by the compiled builder (csrc/plan_builder.cpp, pgb200_plan_build); PGB200_PLAN=python selects the
            # numpy twin (host_setup.build_plan) the tests compare it with
            if self.plan_builder == "python":
                self._plan = build_plan(self._mesh, self._scheme, self._k, self._w, color_fn=_capi.color_cells)
            else:
                self._plan = _capi.plan_build(self._mesh, self._scheme, self.sr, self._k, self._w)
            self._placeholder_k = not self._have_k()
        return self._plan

    def _have_k(self):
        k = self._scheme.k
        return k is not None and np.min(np.abs(k)) >= TOLERANCE

    def _resolve_k(self):
        """response() without k-factors (dcfemmodelling.cpp:1088-1098): analytic ones for flat earth, an error with
        topography"""
        if not getattr(self, "_placeholder_k", False):
            return
        P = self._plan
        if P.topography:
            raise RuntimeError(" data contains no K-factors ")
        self._scheme.k = geometric_factors(self._scheme, self._mesh.dim)
        self.setGeometricFactors(self._scheme.k)

    def setGeometricFactors(self, k):
        """install k-factors (data('k')) without rebuilding the geometry-only plan"""
        k = np.ascontiguousarray(k, np.float64)
        if self._scheme is None or k.size != self._scheme.size:
            raise ValueError("k-factors must have one entry per datum")
        self._scheme.k = k.copy()
        self._placeholder_k = False
        if self._h:
            _capi.check(_capi.lib().pgb200_ert_set_kfac(self._h, self._scheme.k.ctypes.data))

    def _ensure_primary(self):
        """numeric primary potentials with topography (checkPrimpotentials_, dcfemmodelling.cpp:2009-2056): total-field
        solve for rho = 1 on the P2-refined mesh (a second, temporary handle on the same GPU), taken at this mesh's
        nodes -- createP2 keeps the vertex nodes, so the reference's interpolation to mesh_->positions() is a row pick.
        The electrode-potential matrix of that solve also gives the numeric geometric factors (:1539-1556)."""
        P = self._plan
        if not (P.topography and self.sr) or self._prim_pm is not None:
            return
        if self._mesh.order != 1:
            raise NotImplementedError("topography with a P2 secondary mesh: the reference refines with createP2, which "
                                      "needs a P1 mesh")
        prim = CoreB200(sr=False, verbose=self.verbose, device=self.device, preconditioner=self.preconditioner)
        try:
            prim.setMesh(create_p2(self._mesh))
            sch = self._scheme
            prim.setData(SchemeArrays(sch.sensors, sch.a, sch.b, sch.m, sch.n, np.ones(sch.size)))
            prim.setkValues(P.k)
            prim.setWeights(P.w)
            prim.setSolverTolerance(self._tol, self._maxit, self._check)
            P2 = prim._ensure_plan()
            prim.response(np.ones(P2.M))
            ptr, _, _, ld2 = prim._pots_info()
            rows = np.ascontiguousarray(P2.node_inv[P.node_perm], np.int32)   # this handle's node i -> row of the P2 block
            _capi.check(_capi.lib().pgb200_ert_set_primary_dev(self._h, C.c_void_p(ptr), int(ld2), rows.ctypes.data))
            self._prim_pm = prim.get("pm").reshape(P.nE, P.nE)
            self.primary_stats = prim.stats()
        finally:
            prim.close()

    def _ensure_handle(self):
        if self._h is None:
            P = self._ensure_plan()
            h = C.c_void_p()
            native = isinstance(P, _capi.NativePlan)
            if native:
                # device set-up + aggregation hierarchy in the library (pgb200_ert_open_plan)
                rc = _capi.lib().pgb200_ert_open_plan(P._ptr, 1 if self.preconditioner == "multilevel" else 0, self.device, C.byref(h))
                keep = None
            else:
                s, keep = _capi.make_plan_struct(P, self.sr)
                rc = _capi.lib().pgb200_ert_create(C.byref(s), self.device, C.byref(h))
            if rc != 0:
                msg = _capi.last_error()
                if h:
                    _capi.lib().pgb200_ert_destroy(h)
                raise _capi.PGB200Error(msg)
            self._h, self._keep = h, keep
            if self._tol is None:
                self._tol = float(os.environ.get("PGB200_TOL", DEFAULT_PCG_TOL if P.dim == 3 else DEFAULT_PCG_TOL_25D))
            _capi.check(_capi.lib().pgb200_ert_set_solver(h, self._tol, self._maxit, self._check))
            if self._stream:
                _capi.check(_capi.lib().pgb200_ert_set_stream(h, C.c_void_p(self._stream)))
            if self._shard:
                _capi.check(_capi.lib().pgb200_ert_set_shard(h, *self._shard))
            if self._warm:
                _capi.check(_capi.lib().pgb200_ert_set_warm_start(h, 1))
            if native:
                pass
            elif self.preconditioner == "multilevel":
                # aggregation hierarchy from the rho = 1 matrix of the smallest wavenumber (geometry only)
                v1 = self.get("vals1", raw=True).reshape(P.nK, P.nnz)[0]
                self.hierarchy = build_hierarchy(P.rowptr, P.colidx, v1, _capi.pairwise_aggregate)
                self._keep_amg = _capi.set_hierarchy(h, self.hierarchy)
                _capi.check(_capi.lib().pgb200_ert_set_preconditioner(h, 1 if self.hierarchy else 0, 8))
            else:
                _capi.check(_capi.lib().pgb200_ert_set_preconditioner(h, 0, 8))
            if os.environ.get("PGB200_SPMM_VARIANT"):      # A/B switch for measurements: 0 plain gather kernels, 1 streamed
                _capi.check(_capi.lib().pgb200_ert_set_spmm_variant(h, int(os.environ["PGB200_SPMM_VARIANT"])))
            self._ensure_primary()
        return self._h

    # ---- the path ---------------------------------------------------------------------
    def response(self, model):
        if self._complex:
            # response() of a complex fop (dcfemmodelling.cpp:1103-1118): (u_re + i u_im) k, no rounding, no reciprocity mean
            cx = self._ensure_complex()
            m, n_in = self._complex_model(model)
            _capi.check(_capi.lib().pgb200_ert_complex_forward(cx._h, m.ctypes.data, n_in))
            nE = self._scheme.sensors.shape[0]
            pm = cx.get("pm").reshape(2 * nE, 2 * nE)
            u = electrode_matrix_data(pm[:nE, :nE] + 1j * pm[nE:, :nE], self._scheme)
            self._resolve_k_complex()
            resp = u * self._scheme.k
            return np.concatenate([resp.real, resp.imag])
        h = self._ensure_handle()
        self._resolve_k()
        m = np.ascontiguousarray(model, np.float64)
        out = np.zeros(self._scheme.size)
        _capi.check(_capi.lib().pgb200_ert_response(h, m.ctypes.data, int(m.size), out.ctypes.data))
        return out

    def _jacobian_k(self, n_model):
        if self._placeholder_k:
            # prepareJacobianT_ fills missing k-factors first (:1286-1290): analytic, or numeric with topography
            self.setGeometricFactors(self.calcGeometricFactor(nModel=int(n_model)))

    def createJacobian(self, model):
        if self._complex:
            cx = self._ensure_complex()
            m, n_in = self._complex_model(model)
            self._resolve_k_complex()
            D, M = self._scheme.size, self._ensure_plan().M
            J = np.zeros((D, M), np.complex128)
            k = np.ascontiguousarray(self._scheme.k, np.float64)
            _capi.check(_capi.lib().pgb200_ert_complex_jacobian(cx._h, m.ctypes.data, n_in, k.ctypes.data, J.ctypes.data))
            self._Jc = ComplexJacobianB200(J)
            return None
        h = self._ensure_handle()
        self._jacobian_k(np.size(model))
        m = np.ascontiguousarray(model, np.float64)
        _capi.check(_capi.lib().pgb200_ert_create_jacobian(h, m.ctypes.data, int(m.size)))
        return None

    def response_dev(self, model_ptr: int, n: int, out_ptr: int):
        self._ensure_handle()
        self._resolve_k()
        _capi.check(_capi.lib().pgb200_ert_response_dev(self._ensure_handle(), C.c_void_p(model_ptr), int(n), C.c_void_p(out_ptr)))

    def createJacobian_dev(self, model_ptr: int, n: int):
        self._ensure_handle()
        self._jacobian_k(n)
        _capi.check(_capi.lib().pgb200_ert_create_jacobian_dev(self._ensure_handle(), C.c_void_p(model_ptr), int(n)))

    def jacobian(self):
        if self._complex:
            if self._Jc is None:
                raise _capi.PGB200Error("no Jacobian: call createJacobian first")
            return self._Jc
        return self._J

    def _resolve_k_complex(self):
        if self._scheme.k is None:
            if self._ensure_plan().topography:
                raise _capi.PGB200Error(" data contains no K-factors ")
            self._scheme.k = geometric_factors(self._scheme, self._mesh.dim)

    def clearPotentials(self):
        if self._cx is not None and self._cx._h:
            _capi.check(_capi.lib().pgb200_ert_clear_potentials(self._cx._h))
        if self._h:
            _capi.check(_capi.lib().pgb200_ert_clear_potentials(self._h))

    def solution(self):
        """k-summed potentials, one row per electrode (ModellingBase::solution())"""
        P = self._ensure_plan()
        return self.get("solutions").reshape(P.nE, P.N)

    def mapERTModel(self, model, background=-9e99):
        """cell resistivities for a model vector (dcfemmodelling.cpp:1211-1218); only the prolongating
        background (-9e99, the value response() uses) is supported"""
        if background > -9e99:
            raise NotImplementedError("explicit background values are not supported on the B200 path")
        h = self._ensure_handle()
        m = np.ascontiguousarray(model, np.float64)
        out = np.zeros(self._plan.C)
        _capi.check(_capi.lib().pgb200_ert_map_model(h, m.ctypes.data, int(m.size), out.ctypes.data))
        return out

    # ---- generic FEM matrices on the path's element kernels -----------------
    def _fill(self, a, b):
        h = self._ensure_handle()
        P = self._plan
        vec = [None if v is None else np.ascontiguousarray(np.broadcast_to(np.asarray(v, float), (P.C,)), np.float64) for v in (a, b)]
        out = np.zeros(P.nnz)
        _capi.check(_capi.lib().pgb200_ert_fill_matrix(h, *(None if v is None else v.ctypes.data for v in vec), out.ctypes.data))
        return P.ref_rowptr, P.ref_colidx, out[P.ref_slot]

    def fillStiffnessMatrix(self, a=1.0):
        """(rowptr, colidx, vals) of  sum_c a_c int grad N_i . grad N_j  in the reference's CSR layout
        (SparseMatrix::fillStiffnessMatrix, core/src/sparsematrix.h:1034-1049); ``a``: scalar or per-cell"""
        return self._fill(a, None)

    def fillMassMatrix(self, b=1.0):
        """(rowptr, colidx, vals) of  sum_c b_c int N_i N_j  (SparseMatrix::fillMassMatrix, sparsematrix.h:1050-1065)"""
        return self._fill(None, b)

    # ---- introspection ------------------------------------------------------------------
    def get(self, what: str, raw: bool = False) -> np.ndarray:
        h = self._h if (raw and self._h) else self._ensure_handle()
        n = _capi.lib().pgb200_ert_get(h, what.encode(), None, 0)
        if n < 0:
            raise _capi.PGB200Error(_capi.last_error())
        out = np.zeros(int(n))
        if n:
            r = _capi.lib().pgb200_ert_get(h, what.encode(), out.ctypes.data, int(n))
            if r < 0:
                raise _capi.PGB200Error(_capi.last_error())
        if raw:
            return out
        # back to the reference's numbering (the device works in the internal node order)
        P = self._plan
        if what in ("vals", "vals1") and n:
            out = out.reshape(P.nK, P.nnz)[:, P.ref_slot].ravel()
        elif what in ("prim", "pots", "rhs", "sec", "solutions") and n:
            out = out.reshape(-1, P.N)[:, P.node_inv].ravel()
        return out

    def stats(self) -> dict:
        s = np.zeros(17)
        _capi.check(_capi.lib().pgb200_ert_stats(self._ensure_handle(), s.ctypes.data, 17))
        keys = ["pcg_iterations", "max_rel_residual", "launches", "ms_map", "ms_assemble", "ms_rhs", "ms_solve",
                "ms_epilogue", "ms_jacobian", "spmm_timed", "spmm_ms_total", "jacobian_kernel_ms", "jacobian_timed",
                "pcg_iterations_total", "solves", "spmm_bytes_total", "warm_started_solves"]
        return dict(zip(keys, s.tolist()))

    def pathInfo(self) -> dict:
        """which kernels the last solve / Jacobian plan used (pgb200_ert_path_info)"""
        v = np.zeros(10, np.int32)
        _capi.check(_capi.lib().pgb200_ert_path_info(self._ensure_handle(), v.ctypes.data, 10))
        keys = ["spmm_panel_nc", "spmm_tiles", "spmm_two_k", "graph_launches", "jac_chunks", "jac_tiles_per_thread",
                "jac_resolved", "amg_levels", "spmm_slots", "stream_levels"]
        return dict(zip(keys, (int(x) for x in v)))

    def resetStats(self):
        _capi.check(_capi.lib().pgb200_ert_reset_stats(self._ensure_handle()))

    def setProfile(self, on=True):
        """True/1: CUDA events around every SpMM and the Jacobian kernel (no CUDA graph); 2: one event per launch (trace)"""
        _capi.check(_capi.lib().pgb200_ert_set_profile(self._ensure_handle(), int(on)))

    def trace(self):
        """launches since setProfile(2): (source lines in csrc/pgb200_ert.cu, ms since the previous launch finished)"""
        cap = 60000
        lines, ms = np.zeros(cap, np.int32), np.zeros(cap, np.float32)
        n = _capi.lib().pgb200_ert_get_trace(self._ensure_handle(), lines.ctypes.data, ms.ctypes.data, cap)
        if n < 0:
            raise _capi.PGB200Error(_capi.last_error())
        return lines[:n].copy(), ms[:n].copy()

    def _jac_info(self):
        ptr, rows, cols, ld = C.c_void_p(), C.c_int(), C.c_int(), C.c_longlong()
        _capi.check(_capi.lib().pgb200_ert_jacobian_info(self._ensure_handle(), C.byref(ptr), C.byref(rows), C.byref(cols), C.byref(ld)))
        return int(ptr.value or 0), rows.value, cols.value, ld.value

    def _jac_shape(self):
        _, r, c, ld = self._jac_info()
        return r, c, ld

    def _pots_info(self):
        ptr, n, s, ld = C.c_void_p(), C.c_int(), C.c_int(), C.c_longlong()
        _capi.check(_capi.lib().pgb200_ert_potentials_info(self._ensure_handle(), C.byref(ptr), C.byref(n), C.byref(s), C.byref(ld)))
        return int(ptr.value or 0), n.value, s.value, ld.value


class ERTModellingB200:
    """Drop-in for ``pg.physics.ert.ERTModelling`` on the forward + Jacobian path."""

    def __init__(self, sr=True, verbose=False, device=0):
        self._core = CoreB200(sr=sr, verbose=verbose, device=device)
        self._data = None
        # forwarded attributes, as in ertModelling.py:115-120
        self.solution = self._core.solution
        self.calcGeometricFactor = self._core.calcGeometricFactor
        self.mapERTModel = self._core.mapERTModel

    def complex(self):
        return self._core.complex()

    def setComplex(self, c):
        self._core.setComplex(c)

    def setVerbose(self, v):
        self._core.setVerbose(v)

    # data / mesh ------------------------------------------------------------------------
    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, d):
        self.setData(d)

    def setData(self, data):
        self._data = data
        self.setDataPost(data)

    def setDataPost(self, data):
        self._core.setData(data)

    def setMesh(self, mesh, ignoreRegionManager=False):
        self._mesh = mesh
        self.setMeshPost(mesh)

    def setMeshPost(self, mesh):
        self._core.setMesh(mesh, ignoreRegionManager=True)

    def mesh(self):
        return self._mesh

    @property
    def parameterCount(self):
        return int(self._core._ensure_plan().M)

    def createStartModel(self, dataVals):
        return np.full(self.parameterCount, float(np.median(np.asarray(dataVals, float))))

    # the path ---------------------------------------------------------------------------
    def response(self, mod):
        return self._core.response(mod)

    def createJacobian(self, mod):
        if self._core.complex():
            # ertModelling.py:221-236: the complex Jacobian is handed on as the real block matrix [[Re, -Im], [Im, Re]]
            self._core.createJacobian(mod)
            self._Jsq = self._core.jacobian().squeezed(conj=False)
            return self._Jsq
        return self._core.createJacobian(mod)

    def jacobian(self):
        if self._core.complex():
            return self._Jsq
        return self._core.jacobian()
