"""ctypes binding of the C ABI in include/pgb200_ert.h (libpgb200_ert.so, built in-tree).

The product path has no CPU fallback: if the shared library is missing, importing the
compute entry points raises; if it is present but no CUDA device exists, ``create`` fails
with the library's error text.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGB200_LIB") or os.path.join(_PKG, "libpgb200_ert.so")     # PGB200_LIB: A/B builds of the same library

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)


class Plan(C.Structure):
    _fields_ = [
        ("dim", C.c_int), ("nloc", C.c_int), ("n_nodes", C.c_int), ("n_cells", C.c_int), ("nnz", C.c_int),
        ("n_elec", C.c_int), ("n_k", C.c_int), ("n_model", C.c_int), ("n_data", C.c_int), ("sr", C.c_int),
        ("fullspace", C.c_int), ("surface_z", C.c_double),
        ("pos", c_dbl_p), ("cells", c_int_p), ("cell_marker", c_int_p), ("rowptr", c_int_p), ("colidx", c_int_p),
        ("diag_pos", c_int_p),
        ("n_colors", C.c_int), ("color_ptr", c_int_p), ("color_order", c_int_p), ("cells_col", c_int_p), ("pos_col", c_int_p),
        ("k_values", c_dbl_p), ("k_weights", c_dbl_p),
        ("n_bc_slots", C.c_int), ("n_bc_entries", C.c_int), ("bc_slot", c_int_p), ("bc_ptr", c_int_p), ("bc_owner", c_int_p),
        ("bc_coef", c_dbl_p),
        ("n_dir_zero", C.c_int), ("n_dir_nodes", C.c_int), ("dir_zero_slots", c_int_p), ("dir_diag_slots", c_int_p),
        ("dir_nodes", c_int_p),
        ("el_pos", c_dbl_p), ("sing_node", c_int_p), ("sing_val", c_dbl_p), ("pick_ptr", c_int_p), ("pick_idx", c_int_p),
        ("pick_w", c_dbl_p), ("src_cell_ptr", c_int_p), ("src_cells", c_int_p),
        ("n_pro_levels", C.c_int), ("pro_nf", C.c_int), ("pro_level_ptr", c_int_p), ("pro_cells", c_int_p), ("pro_nb", c_int_p),
        ("pro_w", c_dbl_p),
        ("n_jac_cells", C.c_int), ("jac_cells", c_int_p), ("jac_col_ptr", c_int_p),
        ("abmn", c_int_p), ("k_fac", c_dbl_p),
        ("topography", C.c_int), ("ref_node", C.c_int), ("ref_last", C.c_int),
    ]


class MeshIn(C.Structure):
    _fields_ = [("dim", C.c_int), ("nloc", C.c_int), ("n_nodes", C.c_int), ("n_cells", C.c_int), ("n_bounds", C.c_int), ("nlb", C.c_int),
                ("pos", c_dbl_p), ("node_marker", c_int_p), ("cells", c_int_p), ("cell_marker", c_int_p), ("bounds", c_int_p),
                ("bound_marker", c_int_p)]


class SchemeIn(C.Structure):
    _fields_ = [("n_elec", C.c_int), ("n_data", C.c_int), ("sensors", c_dbl_p), ("abmn", c_int_p), ("k_fac", c_dbl_p)]


class AmgLevel(C.Structure):
    _fields_ = [("n", C.c_int), ("nnz", C.c_int), ("rowptr", c_int_p), ("colidx", c_int_p), ("diag_pos", c_int_p),
                ("gal_ptr", c_int_p), ("gal_idx", c_int_p), ("agg", c_int_p), ("mem_ptr", c_int_p), ("mem_idx", c_int_p)]


# every symbol include/pgb200_ert.h declares (checked by tests/test_capi_symbols.py)
EXPORTS = [
    "pgb200_last_error", "pgb200_version", "pgb200_color_cells", "pgb200_build_stream_panels", "pgb200_build_mma_panels", "pgb200_ert_set_spmm_variant", "pgb200_pairwise_aggregate", "pgb200_ert_set_hierarchy", "pgb200_ert_set_preconditioner", "pgb200_ert_set_graph", "pgb200_ert_map_model",
    "pgb200_ert_create", "pgb200_ert_destroy", "pgb200_ert_set_stream", "pgb200_ert_set_solver", "pgb200_ert_set_shard",
    "pgb200_ert_set_kfac", "pgb200_ert_response", "pgb200_ert_create_jacobian", "pgb200_ert_jacobian_copy",
    "pgb200_ert_jacobian_mult", "pgb200_ert_jacobian_tmult", "pgb200_ert_response_dev", "pgb200_ert_create_jacobian_dev",
    "pgb200_ert_jacobian_info", "pgb200_ert_clear_potentials", "pgb200_ert_potentials_info",
    "pgb200_ert_mark_potentials_valid", "pgb200_ert_forward_dev", "pgb200_ert_pm_info", "pgb200_ert_finish_response_dev",
    "pgb200_ert_pack_potentials", "pgb200_ert_get", "pgb200_ert_stats", "pgb200_ert_reset_stats", "pgb200_ert_set_profile",
    "pgb200_spmm", "pgb200_ert_get_trace", "pgb200_ert_set_primary_dev", "pgb200_ert_fill_matrix", "pgb200_ert_jacobian_mult_lr", "pgb200_ert_jacobian_tmult_lr", "pgb200_ert_coverage_trans",
    "pgb200_ert_path_info", "pgb200_ert_bench_spmm", "pgb200_ert_set_complex", "pgb200_ert_complex_forward", "pgb200_ert_complex_jacobian", "pgb200_ert_potentials_state",
    "pgb200_ert_set_warm_start",
    "pgb200_plan_build", "pgb200_plan_free", "pgb200_plan_error", "pgb200_plan_view", "pgb200_plan_array", "pgb200_plan_scalar",
    "pgb200_plan_build_hierarchy", "pgb200_plan_levels", "pgb200_ert_open", "pgb200_ert_open_plan", "pgb200_ert_plan",
]

_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). The B200 ERT path has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.pgb200_last_error.restype = C.c_char_p
        L.pgb200_ert_get.restype = C.c_longlong
        L.pgb200_ert_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]
        L.pgb200_ert_create.argtypes = [C.POINTER(Plan), C.c_int, C.POINTER(C.c_void_p)]
        for name in ("pgb200_ert_destroy", "pgb200_ert_clear_potentials", "pgb200_ert_mark_potentials_valid",
                     "pgb200_ert_reset_stats", "pgb200_ert_potentials_state"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.pgb200_ert_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.pgb200_ert_set_solver.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
        L.pgb200_ert_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.pgb200_ert_set_kfac.argtypes = [C.c_void_p, C.c_void_p]
        L.pgb200_ert_response.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.pgb200_ert_response_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.pgb200_ert_create_jacobian.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.pgb200_ert_create_jacobian_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.pgb200_ert_jacobian_copy.argtypes = [C.c_void_p, C.c_void_p]
        L.pgb200_ert_jacobian_mult.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pgb200_ert_jacobian_tmult.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pgb200_ert_jacobian_mult_lr.argtypes = [C.c_void_p] * 5
        L.pgb200_ert_jacobian_tmult_lr.argtypes = [C.c_void_p] * 5
        L.pgb200_ert_coverage_trans.argtypes = [C.c_void_p] * 4
        L.pgb200_ert_get_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.pgb200_ert_fill_matrix.argtypes = [C.c_void_p] * 4
        L.pgb200_ert_set_primary_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
        L.pgb200_ert_jacobian_info.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), c_int_p, c_int_p, C.POINTER(C.c_longlong)]
        L.pgb200_ert_potentials_info.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), c_int_p, c_int_p, C.POINTER(C.c_longlong)]
        L.pgb200_ert_forward_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.pgb200_ert_pm_info.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), c_int_p]
        L.pgb200_ert_finish_response_dev.argtypes = [C.c_void_p, C.c_void_p]
        L.pgb200_ert_pack_potentials.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.pgb200_ert_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.pgb200_ert_set_profile.argtypes = [C.c_void_p, C.c_int]
        L.pgb200_ert_path_info.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.pgb200_ert_bench_spmm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.pgb200_ert_set_complex.argtypes = [C.c_void_p, C.c_int]
        L.pgb200_ert_complex_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.pgb200_ert_complex_jacobian.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.pgb200_plan_error.restype = C.c_char_p
        L.pgb200_plan_build.argtypes = [C.POINTER(MeshIn), C.POINTER(SchemeIn), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.pgb200_plan_free.argtypes = [C.c_void_p]
        L.pgb200_plan_view.argtypes = [C.c_void_p]
        L.pgb200_plan_view.restype = C.POINTER(Plan)
        L.pgb200_plan_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), c_int_p]
        L.pgb200_plan_scalar.argtypes = [C.c_void_p, C.c_char_p, c_dbl_p]
        L.pgb200_plan_build_hierarchy.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]
        L.pgb200_plan_levels.argtypes = [C.c_void_p]
        L.pgb200_plan_levels.restype = C.c_void_p
        L.pgb200_ert_open.argtypes = [C.POINTER(MeshIn), C.POINTER(SchemeIn), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.pgb200_ert_open_plan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.pgb200_ert_plan.argtypes = [C.c_void_p]
        L.pgb200_ert_plan.restype = C.c_void_p
        L.pgb200_color_cells.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.pgb200_build_stream_panels.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 11
        L.pgb200_build_mma_panels.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 9
        L.pgb200_ert_set_spmm_variant.argtypes = [C.c_void_p, C.c_int]
        L.pgb200_ert_set_hierarchy.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.pgb200_ert_set_preconditioner.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.pgb200_ert_set_graph.argtypes = [C.c_void_p, C.c_int]
        L.pgb200_ert_set_warm_start.argtypes = [C.c_void_p, C.c_int]
        L.pgb200_ert_map_model.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.pgb200_pairwise_aggregate.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.pgb200_spmm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p]
        _lib = L
    return _lib


def last_error() -> str:
    return lib().pgb200_last_error().decode()


class PGB200Error(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise PGB200Error(last_error())


def color_cells(cells: np.ndarray, n_nodes: int):
    """C++ greedy colouring (host helper of the library, no GPU needed)."""
    cells = np.ascontiguousarray(cells, np.int32)
    color = np.zeros(cells.shape[0], np.int32)
    n = lib().pgb200_color_cells(cells.shape[0], cells.shape[1], cells.ctypes.data, int(n_nodes), color.ctypes.data)
    if n <= 0:
        raise PGB200Error(last_error())
    return color, int(n)


AGGREGATION_THETA = 0.25     # strength threshold of the pairwise matching (see pgb200_pairwise_aggregate)


def pairwise_aggregate(rowptr, colidx, vals, group=None, theta=None):
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    colidx = np.ascontiguousarray(colidx, np.int32)
    vals = np.ascontiguousarray(vals, np.float64)
    agg = np.zeros(rowptr.size - 1, np.int32)
    g = None if group is None else np.ascontiguousarray(group, np.int32)
    na = lib().pgb200_pairwise_aggregate(rowptr.size - 1, rowptr.ctypes.data, colidx.ctypes.data, vals.ctypes.data,
                                         None if g is None else g.ctypes.data,
                                         C.c_double(AGGREGATION_THETA if theta is None else float(theta)), agg.ctypes.data)
    if na < 0:
        raise PGB200Error(last_error())
    return agg, int(na)


def build_mma_panels(rowptr: np.ndarray, colidx: np.ndarray, groups: int = 12, hc: int = 104, max_chunks: int = 8, rowb_hint: int = 800) -> dict:
    """8-row-group form of the streamed panels (k_spmm_mma, csrc/stream_panels.h) -- host-side tests"""
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    colidx = np.ascontiguousarray(colidx, np.int32)
    n = rowptr.size - 1
    counts = np.zeros(8, np.int32)
    args = [n, rowptr.ctypes.data, colidx.ctypes.data, int(groups), int(hc), int(max_chunks), int(rowb_hint), counts.ctypes.data]
    if lib().pgb200_build_mma_panels(*args, *([None] * 8)) != 0:
        raise PGB200Error(last_error())
    npan, nch, nh, nks, nmeta, gstride, mks, mmeta = (int(x) for x in counts)
    out = dict(panel_row_ptr=np.zeros(npan + 1, np.int32), panel_chunk_ptr=np.zeros(npan + 1, np.int32),
               chunk_halo_ptr=np.zeros(nch + 1, np.int32), halo_cols=np.zeros(max(1, nh), np.int32),
               chunk_ks_ptr=np.zeros(nch + 1, np.int32), a_src=np.zeros(max(1, 32 * nks), np.int32),
               chunk_meta_ptr=np.zeros(nch + 1, np.int32), meta=np.zeros(max(1, nmeta), np.uint32))
    if lib().pgb200_build_mma_panels(*args, *(out[k].ctypes.data for k in ("panel_row_ptr", "panel_chunk_ptr", "chunk_halo_ptr", "halo_cols",
                                                                          "chunk_ks_ptr", "a_src", "chunk_meta_ptr", "meta"))) != 0:
        raise PGB200Error(last_error())
    out.update(n_panels=npan, n_chunks=nch, n_ks=nks, meta_gstride=gstride, max_chunk_ks=mks, max_chunk_meta=mmeta, groups=int(groups))
    return out


def build_stream_panels(rowptr: np.ndarray, colidx: np.ndarray, rmax: int = 64, hc: int = 104, max_chunks: int = 2) -> dict:
    """streamed row panels of a CSR pattern (csrc/stream_panels.h); the library builds them itself, this wrapper serves
    the host-side tests that replay the kernel's traversal"""
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    colidx = np.ascontiguousarray(colidx, np.int32)
    n = rowptr.size - 1
    counts = np.zeros(10, np.int32)
    args = [n, rowptr.ctypes.data, colidx.ctypes.data, int(rmax), int(hc), int(max_chunks), counts.ctypes.data]
    if lib().pgb200_build_stream_panels(*args, *([None] * 10)) != 0:
        raise PGB200Error(last_error())
    npan, nch, nh, stride, max_rows, mch, mce, nnz, nruns, mruns = (int(x) for x in counts)
    out = dict(panel_row_ptr=np.zeros(npan + 1, np.int32), panel_chunk_ptr=np.zeros(npan + 1, np.int32),
               chunk_halo_ptr=np.zeros(nch + 1, np.int32), halo_cols=np.zeros(max(1, nh), np.int32),
               chunk_ent_ptr=np.zeros(nch + 1, np.int32), ent_src=np.zeros(max(1, nnz), np.int32),
               ent_idx=np.zeros(max(1, nnz), np.uint32), crp=np.zeros(max(1, nch * stride), np.int32),
               chunk_run_ptr=np.zeros(nch + 1, np.int32), runs=np.zeros((max(1, nruns), 3), np.int32))
    if lib().pgb200_build_stream_panels(*args, *(out[k].ctypes.data for k in ("panel_row_ptr", "panel_chunk_ptr", "chunk_halo_ptr",
                                                                             "halo_cols", "chunk_ent_ptr", "ent_src", "ent_idx", "crp",
                                                                             "chunk_run_ptr", "runs"))) != 0:
        raise PGB200Error(last_error())
    out.update(n_runs=nruns, max_chunk_runs=mruns)
    out.update(n_panels=npan, n_chunks=nch, crp_stride=stride, max_rows=max_rows, max_chunk_halo=mch, max_chunk_ent=mce, nnz=nnz)
    return out


def set_hierarchy(handle, levels):
    """levels: output of amg_setup.build_hierarchy"""
    keep = []
    arr = (AmgLevel * max(1, len(levels)))()
    for i, L in enumerate(levels):
        a = arr[i]
        a.n, a.nnz = int(L["n"]), int(L["nnz"])
        for name in ("rowptr", "colidx", "diag_pos", "gal_ptr", "gal_idx", "agg", "mem_ptr", "mem_idx"):
            v = np.ascontiguousarray(L[name], np.int32)
            keep.append(v)
            setattr(a, name, v.ctypes.data_as(c_int_p))
    check(lib().pgb200_ert_set_hierarchy(handle, len(levels), C.cast(arr, C.c_void_p)))
    return keep


class NativePlan:
    """A plan built by the compiled builder (csrc/plan_builder.cpp, pgb200_plan_build).  Attribute access mirrors
    host_setup.ERTPlan for everything the Python layer reads: sizes (N C nnz nE nK nS M dim nloc), flags (topography
    has_background), and arrays by name (k w rowptr colidx ref_rowptr ref_colidx ref_slot node_perm node_inv ...), fetched
    lazily as numpy copies."""

    _SCALARS = dict(N="N", C="C", nnz="nnz", nE="nE", nK="nK", M="M", dim="dim", nloc="nloc", n_colors="n_colors", pro_nf="pro_nf")

    def __init__(self, ptr, mesh, scheme, keep):
        self._ptr, self.mesh_ref, self.scheme, self._keep, self._cache = ptr, mesh, scheme, keep, {}

    def free(self):
        if self._ptr:
            lib().pgb200_plan_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def scalar(self, name):
        v = C.c_double()
        if lib().pgb200_plan_scalar(self._ptr, name.encode(), C.byref(v)) != 0:
            raise PGB200Error(lib().pgb200_plan_error().decode())
        return v.value

    def array(self, name):
        if name in self._cache:
            return self._cache[name]
        ptr, n, t = C.c_void_p(), C.c_longlong(), C.c_int()
        if lib().pgb200_plan_array(self._ptr, name.encode(), C.byref(ptr), C.byref(n), C.byref(t)) != 0:
            raise AttributeError(lib().pgb200_plan_error().decode())
        dt = {0: np.int32, 1: np.float64, 2: np.int64}[t.value]
        if n.value == 0:
            out = np.zeros(0, dt)
        else:
            out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER({0: C.c_int, 1: C.c_double, 2: C.c_longlong}[t.value])), shape=(n.value,)).copy()
        self._cache[name] = out
        return out

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if name in NativePlan._SCALARS:
            return int(self.scalar(NativePlan._SCALARS[name]))
        if name == "nS":
            return int(self.scalar("nK")) * int(self.scalar("nE"))
        if name in ("topography", "has_background", "neumann_domain", "k_missing"):
            return bool(self.scalar(name))
        if name in ("ref_node", "ref_last"):
            return int(self.scalar(name))
        if name == "surface_z":
            return self.scalar("surface_z")
        if name == "dir_zero_slots" or name == "dir_diag_slots":
            return self.array(name)
        return self.array(name)

    def view(self):
        return lib().pgb200_plan_view(self._ptr)


def plan_build(mesh, scheme, sr=True, k_values=None, weights=None) -> NativePlan:
    """mesh (MeshArrays) + scheme (SchemeArrays) -> NativePlan through pgb200_plan_build (compiled, OpenMP)"""
    keep = [np.ascontiguousarray(mesh.pos, np.float64), np.ascontiguousarray(mesh.node_marker, np.int32),
            np.ascontiguousarray(mesh.cells, np.int32), np.ascontiguousarray(mesh.cell_marker, np.int32),
            np.ascontiguousarray(mesh.bounds, np.int32).reshape(mesh.bound_marker.size, -1),
            np.ascontiguousarray(mesh.bound_marker, np.int32),
            np.ascontiguousarray(scheme.sensors, np.float64), np.ascontiguousarray(scheme.abmn(), np.int32)]
    m = MeshIn()
    m.dim, m.nloc, m.n_nodes, m.n_cells = mesh.dim, mesh.cells.shape[1], mesh.pos.shape[0], mesh.cells.shape[0]
    m.n_bounds, m.nlb = keep[5].size, (keep[4].shape[1] if keep[4].size else mesh.dim)
    m.pos, m.node_marker, m.cells, m.cell_marker, m.bounds, m.bound_marker = _dp(keep[0]), _ip(keep[1]), _ip(keep[2]), _ip(keep[3]), _ip(keep[4]), _ip(keep[5])
    sc = SchemeIn()
    sc.n_elec, sc.n_data, sc.sensors, sc.abmn = keep[6].shape[0], keep[7].shape[0], _dp(keep[6]), _ip(keep[7])
    if scheme.k is not None:
        kf = np.ascontiguousarray(scheme.k, np.float64)
        keep.append(kf)
        sc.k_fac = _dp(kf)
    nk, kp, wp = 0, None, None
    if k_values is not None and weights is not None:
        ka, wa = np.ascontiguousarray(k_values, np.float64), np.ascontiguousarray(weights, np.float64)
        keep.extend([ka, wa])
        nk, kp, wp = ka.size, ka.ctypes.data, wa.ctypes.data
    out = C.c_void_p()
    if lib().pgb200_plan_build(C.byref(m), C.byref(sc), 1 if sr else 0, nk, kp, wp, C.byref(out)) != 0:
        msg = lib().pgb200_plan_error().decode()
        if "not on the B200 path" in msg or "not supported" in msg:
            raise NotImplementedError(msg)
        if "does not match the given mesh" in msg:
            raise ValueError(msg)
        raise PGB200Error(msg)
    return NativePlan(out, mesh, scheme, keep)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _dp(a):
    return a.ctypes.data_as(c_dbl_p)


def make_plan_struct(P, sr: bool):
    """ERTPlan (host_setup.build_plan) -> (ctypes Plan, keep-alive list of arrays)."""
    keep = []

    def I(a):
        a = np.ascontiguousarray(a, np.int32)
        keep.append(a)
        return _ip(a)

    def D(a):
        a = np.ascontiguousarray(a, np.float64)
        keep.append(a)
        return _dp(a)

    s = Plan()
    s.dim, s.nloc, s.n_nodes, s.n_cells, s.nnz = P.dim, P.nloc, P.N, P.C, P.nnz
    s.n_elec, s.n_k, s.n_model, s.n_data, s.sr = P.nE, P.nK, P.M, P.scheme.size, 1 if sr else 0
    s.fullspace = 1 if P.surface_z <= -1e300 else 0
    s.surface_z = 0.0 if s.fullspace else P.surface_z
    s.topography = 1 if getattr(P, "topography", False) else 0
    s.ref_node, s.ref_last = int(getattr(P, "ref_node", -1)), int(getattr(P, "ref_last", 0))
    s.pos, s.cells, s.cell_marker = D(P.mesh.pos), I(P.mesh.cells), I(P.cell_marker)
    s.rowptr, s.colidx, s.diag_pos = I(P.rowptr), I(P.colidx), I(P.diag_pos)
    s.n_colors, s.color_ptr, s.color_order = P.n_colors, I(P.color_ptr), I(P.color_order)
    s.cells_col, s.pos_col = I(P.cells_col), I(P.pos_col)
    s.k_values, s.k_weights = D(P.k), D(P.w)
    s.n_bc_slots, s.n_bc_entries = int(P.bc_slot.size), int(P.bc_owner.size)
    s.bc_slot, s.bc_ptr, s.bc_owner, s.bc_coef = I(P.bc_slot), I(P.bc_ptr), I(P.bc_owner), D(P.bc_coef)
    s.n_dir_zero, s.n_dir_nodes = int(P.dir_zero_slots.size), int(P.dir_nodes.size)
    s.dir_zero_slots, s.dir_diag_slots, s.dir_nodes = I(P.dir_zero_slots), I(P.dir_diag_slots), I(P.dir_nodes)
    s.el_pos, s.sing_node, s.sing_val = D(P.el_pos), I(P.sing_node), D(P.sing_val)
    s.pick_ptr, s.pick_idx, s.pick_w = I(P.pick_ptr), I(P.pick_idx), D(P.pick_w)
    s.src_cell_ptr, s.src_cells = I(P.src_cell_ptr), I(P.src_cells)
    lv = P.pro_levels
    s.n_pro_levels = len(lv)
    s.pro_nf = int(getattr(P, "pro_nf", 0)) if lv else 0
    lp = np.concatenate([[0], np.cumsum([len(c) for c, _, _ in lv])]).astype(np.int32) if lv else np.zeros(1, np.int32)
    s.pro_level_ptr = I(lp)
    s.pro_cells = I(np.concatenate([c for c, _, _ in lv]) if lv else np.zeros(0, np.int32))
    s.pro_nb = I(np.concatenate([n for _, n, _ in lv]).ravel() if lv else np.zeros(0, np.int32))
    s.pro_w = D(np.concatenate([w for _, _, w in lv]).ravel() if lv else np.zeros(0))
    s.n_jac_cells, s.jac_cells, s.jac_col_ptr = int(P.jac_cells.size), I(P.jac_cells), I(P.jac_col_ptr)
    s.abmn = I(P.scheme.abmn())
    kf = P.scheme.k if P.scheme.k is not None else np.zeros(P.scheme.size)
    s.k_fac = D(kf)
    return s, keep
