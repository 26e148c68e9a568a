"""Run under torchrun with N ranks: sharded forward + Jacobian must equal the single-GPU result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from cases import make_case  # noqa: E402
from pygimli_b200.dist import ShardedERT  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name in ("2d_p1", "3d_p1"):
        mesh, scheme, model = make_case(name)
        single = ShardedERT(mesh, scheme, device=local, rank=0, world=1)
        r1 = single.response(model)
        single.create_jacobian(model)
        J1 = single.core.jacobian().numpy()
        x = np.random.default_rng(1).standard_normal(J1.shape[1])
        y = np.random.default_rng(2).standard_normal(J1.shape[0])
        sh = ShardedERT(mesh, scheme, device=local, rank=rank, world=world)
        r2 = sh.response(model)
        sh.create_jacobian(model)
        Jloc, rows = sh.jacobian_rows()
        e_r = np.max(np.abs(r2 - r1) / np.abs(r1))
        e_J = np.max(np.abs(Jloc - J1[rows])) / np.max(np.abs(J1)) if rows.size else 0.0
        e_x = np.max(np.abs(sh.jac_mult(x) - J1 @ x)) / np.max(np.abs(J1 @ x))
        e_y = np.max(np.abs(sh.jac_tmult(y) - J1.T @ y)) / np.max(np.abs(J1.T @ y))
        dd, mm = 1.0 / r1, 1.0 / model
        cov1 = np.abs(J1 * dd[:, None]).sum(0) / np.abs(mm)
        e_c = np.max(np.abs(sh.coverage_trans(dd, mm) - cov1)) / np.max(cov1)
        good = e_r < 1e-9 and e_J < 1e-9 and e_x < 1e-9 and e_y < 1e-9 and e_c < 1e-9
        ok = ok and good
        print(f"rank {rank} {name}: rows {rows.size} sources {sh.n_local_sources}  rhoa {e_r:.2e}  J {e_J:.2e}  Jx {e_x:.2e}  JTy {e_y:.2e}  cov {e_c:.2e}  {'OK' if good else 'FAIL'}", flush=True)
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if t.item() < 1.0:
        sys.exit(1)


if __name__ == "__main__":
    main()
