"""One-process-per-GPU sharding of the ERT path.

What shards, and what is exchanged:
  * solve phase: the nS = nE x nK current sources are independent given the (replicated,
    cheap to assemble) matrices -> rank r solves the contiguous source range src_range(r);
    no communication while iterating.
  * forward response: only the nE x nE electrode-potential matrix is needed -> every rank
    sums its own sources into a partial matrix, one NCCL all-reduce (nE^2 doubles).
  * Jacobian: data rows are independent -> rank r writes rows row_range(r) (rows are ordered
    by current dipole so a shard touches few current electrodes).  A row needs the potentials
    of a, b, m and n, so the k-resolved potentials are exchanged ONCE per iteration with an
    NCCL all-gather (N x nS doubles in total); everything else stays local.
  * J.x needs an all-gather of D doubles, J^T.y and the coverage (column sums of |J|) an all-reduce of M doubles.
torch.distributed is plumbing only; all arithmetic is in libpgb200_ert.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .ert_modelling import CoreB200
from .scheme import SchemeArrays


def split_range(n: int, world: int, rank: int):
    """contiguous, balanced partition of range(n)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def chunk_width(n: int, world: int) -> int:
    return -(-n // world)


def padded_range(n: int, world: int, rank: int):
    """balanced contiguous chunks whose boundaries are EVEN (the panel-staged SpMM needs 16-byte aligned
    column tiles); the all-gather uses slots of the largest chunk width"""
    def bound(r):
        return n if r >= world else min(n, 2 * int(round(r * n / (2.0 * world))))
    return bound(rank), bound(rank + 1)


def max_chunk(n: int, world: int) -> int:
    return max(padded_range(n, world, r)[1] - padded_range(n, world, r)[0] for r in range(world))


def row_order(scheme: SchemeArrays) -> np.ndarray:
    """data rows ordered by current dipole (a, b): a contiguous shard then needs only a few
    current-side electrodes in its Gram block"""
    return np.lexsort((scheme.b, scheme.a)).astype(np.int64)


class _DevView:
    """expose a raw device pointer to torch (zero-copy) through the CUDA array interface"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = dict(shape=(n,), typestr="<f8", data=(ptr, False), version=3, strides=None)


class ShardedERT:
    def __init__(self, mesh, scheme: SchemeArrays, device=0, rank=0, world=1, sr=True, preconditioner="multilevel", kw=None):
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        self.perm = row_order(scheme) if world > 1 else np.arange(scheme.size)
        self.inv_perm = np.argsort(self.perm)
        self.scheme = scheme.subset(self.perm) if world > 1 else scheme
        self.core = CoreB200(sr=sr, device=device, preconditioner=preconditioner)
        self.core.setMesh(mesh)
        self.core.setData(self.scheme)
        if kw is not None:
            self.core.setkValues(kw[0])
            self.core.setWeights(kw[1])
        P = self.core._ensure_plan()
        self.nS, self.N, self.nE, self.D, self.M = P.nS, P.N, P.nE, self.scheme.size, P.M
        self.src = padded_range(self.nS, world, rank) if world > 1 else (0, self.nS)
        self.rows = split_range(self.D, world, rank) if world > 1 else (0, self.D)
        if world > 1:
            self.core.setShard(self.src[0], self.src[1], self.rows[0], self.rows[1])
        self.core._ensure_handle()      # with topography every rank solves the (replicated) P2 primary problem here
        self.core._resolve_k()
        self._bufs = None

    @property
    def n_local_sources(self) -> int:
        return self.src[1] - self.src[0]

    def set_solver(self, tol, max_iter, check_every):
        self.core.setSolverTolerance(tol, max_iter, check_every)

    def set_stream(self, stream_ptr):
        self.core.setStream(stream_ptr)

    # ---- collectives ------------------------------------------------------------------
    def _torch(self):
        import torch
        import torch.distributed as dist
        return torch, dist

    def _allreduce_pm(self):
        torch, dist = self._torch()
        ptr, n = C.c_void_p(), C.c_int()
        _capi.check(_capi.lib().pgb200_ert_pm_info(self.core._h, C.byref(ptr), C.byref(n)))
        t = torch.as_tensor(_DevView(ptr.value, n.value), device=f"cuda:{self.device}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def _allgather_potentials(self):
        torch, dist = self._torch()
        w = max_chunk(self.nS, self.world)
        if self._bufs is None:
            self._bufs = (torch.zeros(self.N * w, dtype=torch.float64, device=f"cuda:{self.device}"),
                          torch.zeros(self.world * self.N * w, dtype=torch.float64, device=f"cuda:{self.device}"))
        send, recv = self._bufs
        L = _capi.lib()
        c0, c1 = self.src
        if c1 > c0:
            # pack as [N x (c1-c0)] at the start of the fixed-width slot
            _capi.check(L.pgb200_ert_pack_potentials(self.core._h, c0, c1, C.c_void_p(send.data_ptr()), 0))
        dist.all_gather_into_tensor(recv, send)
        for r in range(self.world):
            if r == self.rank:
                continue
            a, b = padded_range(self.nS, self.world, r)
            if b > a:
                _capi.check(L.pgb200_ert_pack_potentials(self.core._h, a, b, C.c_void_p(recv.data_ptr() + 8 * r * self.N * w), 1))
        _capi.check(L.pgb200_ert_mark_potentials_valid(self.core._h))

    # ---- device-resident path ----------------------------------------------------------
    def response_dev(self, model_dev, rhoa_dev):
        """rhoa_dev receives the apparent resistivities in the ORIGINAL data order"""
        if self.world == 1:
            self.core.response_dev(model_dev.data_ptr(), model_dev.numel(), rhoa_dev.data_ptr())
            return
        torch, _ = self._torch()
        L = _capi.lib()
        _capi.check(L.pgb200_ert_forward_dev(self.core._h, C.c_void_p(model_dev.data_ptr()), int(model_dev.numel())))
        self._allreduce_pm()
        tmp = torch.empty(self.D, dtype=torch.float64, device=rhoa_dev.device)
        _capi.check(L.pgb200_ert_finish_response_dev(self.core._h, C.c_void_p(tmp.data_ptr())))
        rhoa_dev.copy_(tmp[torch.as_tensor(self.inv_perm, device=rhoa_dev.device)])

    def create_jacobian_dev(self, model_dev):
        if self.world > 1:
            state = _capi.lib().pgb200_ert_potentials_state(self.core._h)
            if not (state & 2):
                # prepareJacobianT_ (dcfemmodelling.cpp:1246-1309): no potentials (never solved, setShard, clearPotentials)
                # -> solve this shard's sources for this model first; only then are the other shards' columns gathered
                if not (state & 1):
                    _capi.check(_capi.lib().pgb200_ert_forward_dev(self.core._h, C.c_void_p(model_dev.data_ptr()), int(model_dev.numel())))
                self._allgather_potentials()
        self.core.createJacobian_dev(model_dev.data_ptr(), model_dev.numel())

    # ---- host-buffer path ----------------------------------------------------------------
    def response(self, model: np.ndarray) -> np.ndarray:
        if self.world == 1:
            return self.core.response(model)
        torch, _ = self._torch()
        md = torch.from_numpy(np.ascontiguousarray(model, np.float64)).to(f"cuda:{self.device}")
        out = torch.empty(self.D, dtype=torch.float64, device=md.device)
        self.response_dev(md, out)
        return out.cpu().numpy()

    def create_jacobian(self, model: np.ndarray):
        if self.world == 1:
            return self.core.createJacobian(model)
        torch, _ = self._torch()
        md = torch.from_numpy(np.ascontiguousarray(model, np.float64)).to(f"cuda:{self.device}")
        self.create_jacobian_dev(md)

    def jac_mult(self, x: np.ndarray) -> np.ndarray:
        y_loc = self.core.jacobian().mult(x)
        if self.world == 1:
            return y_loc
        torch, dist = self._torch()
        w = chunk_width(self.D, self.world) + 1
        send = torch.zeros(w, dtype=torch.float64, device=f"cuda:{self.device}")
        send[: y_loc.size] = torch.from_numpy(y_loc).to(send.device)
        recv = torch.zeros(self.world * w, dtype=torch.float64, device=send.device)
        dist.all_gather_into_tensor(recv, send)
        recv = recv.cpu().numpy()
        y = np.zeros(self.D)
        for r in range(self.world):
            a, b = split_range(self.D, self.world, r)
            y[a:b] = recv[r * w: r * w + (b - a)]
        return y[self.inv_perm]

    def jac_tmult(self, y: np.ndarray) -> np.ndarray:
        if self.world == 1:
            return self.core.jacobian().transMult(y)
        torch, dist = self._torch()
        yp = np.asarray(y, float)[self.perm][self.rows[0]: self.rows[1]]
        x_loc = self.core.jacobian().transMult(yp) if yp.size else np.zeros(self.M)
        t = torch.from_numpy(x_loc).to(f"cuda:{self.device}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def coverage_trans(self, dd: np.ndarray, mm: np.ndarray) -> np.ndarray:
        """coverageDCtrans (bertJacobian.cpp:569) over the row-sharded J: local column sums of |J_ij dd_i|, one
        all-reduce of M doubles, then the division by |mm|"""
        if self.world == 1:
            return self.core.jacobian().coverageDCtrans(dd, mm)
        torch, dist = self._torch()
        ddp = np.asarray(dd, float)[self.perm][self.rows[0]: self.rows[1]]
        part = self.core.jacobian().coverageDCtrans(ddp) if ddp.size else np.zeros(self.M)
        t = torch.from_numpy(part).to(f"cuda:{self.device}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy() / np.abs(np.asarray(mm, float))

    def jacobian_rows(self) -> np.ndarray:
        """this rank's rows of J (row-major), and their ORIGINAL data indices"""
        return self.core.jacobian().numpy(), self.perm[self.rows[0]: self.rows[1]]
