"""The thin PyTorch C++ extension of the B200 ERT path (csrc/torch_ext.cpp -> libpgb200_torch.so, ``torch.ops.pgb200``).

``TorchERT`` is the tensor-in / tensor-out door: mesh and scheme go to the compiled plan builder once
(``pgb200_ert_open``), ``response(model)`` and ``create_jacobian(model)`` take float64 CUDA tensors that are already
resident in HBM and run on torch's current stream; the Jacobian comes back as a zero-copy (rows, cols) view of the
HBM-resident buffer.  Mirrors ``ERTModelling.response`` / ``createJacobian`` (pygimli/physics/ert/ertModelling.py:213, :238).
There is no fallback: without the built extension or without a CUDA device the calls raise.
"""
from __future__ import annotations

import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
TORCH_LIB_PATH = os.path.join(_PKG, "libpgb200_torch.so")
_loaded = False


def ops():
    global _loaded
    import torch
    if not _loaded:
        if not os.path.exists(TORCH_LIB_PATH):
            raise RuntimeError(f"{TORCH_LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        torch.ops.load_library(TORCH_LIB_PATH)
        _loaded = True
    return torch.ops.pgb200


class TorchERT:
    def __init__(self, mesh, scheme, sr: bool = True, multilevel: bool = True, device: int = 0):
        import torch
        t = torch.from_numpy
        k = None if scheme.k is None else t(np.ascontiguousarray(scheme.k, np.float64))
        bounds = np.ascontiguousarray(mesh.bounds, np.int32).reshape(mesh.bound_marker.size, -1)
        self.device = int(device)
        self.D, self.M = int(scheme.size), int(mesh.cell_marker.max()) + 1
        self._h = ops().open(t(np.ascontiguousarray(mesh.pos, np.float64)), t(np.ascontiguousarray(mesh.node_marker, np.int32)),
                             t(np.ascontiguousarray(mesh.cells, np.int32)), t(np.ascontiguousarray(mesh.cell_marker, np.int32)),
                             t(bounds), t(np.ascontiguousarray(mesh.bound_marker, np.int32)),
                             t(np.ascontiguousarray(scheme.sensors, np.float64)), t(np.ascontiguousarray(scheme.abmn(), np.int32)),
                             k, int(mesh.dim), bool(sr), bool(multilevel), self.device)

    def response(self, model):
        """model: float64 CUDA tensor [M] (or [C]) -> apparent resistivities, float64 CUDA tensor [D]"""
        return ops().response(self._h, model)

    def create_jacobian(self, model):
        """-> zero-copy (D, M) view of the HBM-resident Jacobian (valid until the next create_jacobian / close)"""
        return ops().create_jacobian(self._h, model)

    def jac_mult(self, x):
        return ops().jac_mult(self._h, x)

    def jac_tmult(self, y):
        return ops().jac_tmult(self._h, y)

    def set_solver(self, tol=1e-12, max_iter=50000, check_every=25):
        ops().set_solver(self._h, float(tol), int(max_iter), int(check_every))

    def stats(self):
        return ops().stats(self._h)

    def close(self):
        if self._h:
            ops().close(self._h)
            self._h = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
