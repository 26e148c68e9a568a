"""The thin PyTorch extension (csrc/torch_ext.cpp, torch.ops.pgb200): registration and loud failure on the CPU,
parity with the golden vectors of the reference through the tensor-in / tensor-out door on the GPU."""
import os

import numpy as np
import pytest

from cases import make_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_ops_are_registered_and_fail_loudly_without_gpu():
    import torch
    from pygimli_b200 import torch_ops
    o = torch_ops.ops()
    for name in ("open", "response", "create_jacobian", "jac_mult", "jac_tmult", "set_solver", "stats", "close"):
        assert hasattr(o, name)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    mesh, scheme, _ = make_case("2d_p1")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        torch_ops.TorchERT(mesh, scheme)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["2d_p1", "3d_p1"])
def test_torch_door_matches_reference(name):
    import torch
    from pygimli_b200 import torch_ops
    mesh, scheme, model = make_case(name)
    g = np.load(os.path.join(GOLD, name + ".npz"))
    fop = torch_ops.TorchERT(mesh, scheme)
    m = torch.from_numpy(model).cuda()
    rhoa = fop.response(m).cpu().numpy()
    kf = np.abs(g["kfac"])
    assert np.all(np.abs(rhoa - g["rhoa"]) <= 1e-8 * np.abs(g["rhoa"]) + 2e-10 * kf)
    J = fop.create_jacobian(m)
    assert tuple(J.shape) == g["J"].shape and J.is_cuda
    Jh = J.cpu().numpy()
    assert np.max(np.abs(Jh - g["J"])) / np.max(np.abs(g["J"])) < 1e-8
    x = np.random.default_rng(0).standard_normal(Jh.shape[1])
    assert np.max(np.abs(fop.jac_mult(torch.from_numpy(x)).numpy() - Jh @ x)) <= 1e-12 * np.max(np.abs(Jh @ x))
    assert fop.stats()[1].item() <= 1e-12 * 1.0001
    fop.close()
