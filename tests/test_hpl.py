"""HPLinpack retarget (HPL-MxP): CPU (stub transport) at world 1 and 2; GPU variants use the tcgen05 GEMM and NVLink broadcasts."""
import pytest
import torch

from _mp import run_ranks


@pytest.mark.parametrize("world,n,nb", [(1, 512, 64), (2, 512, 64), (2, 384, 128)])
def test_hpl_mxp_cpu_stub(world, n, nb):
    ok, outs = run_ranks("_hpl_worker.py", world, extra=["--n", str(n), "--nb", str(nb)], timeout=280)
    assert ok, "\n".join(outs)


def test_lu_nopivot_block_matches_dense():
    from batch_shipyard_b200.models.hpl import _lu_nopivot_
    torch.manual_seed(0)
    a = torch.rand(96, 96, dtype=torch.float64) - 0.5 + 48 * torch.eye(96, dtype=torch.float64)
    d = a.clone()
    _lu_nopivot_(d, base=8)
    lo = torch.tril(d, -1) + torch.eye(96, dtype=torch.float64)
    up = torch.triu(d)
    assert float((lo @ up - a).abs().max()) < 1e-12


@pytest.mark.gpu
def test_hpl_mxp_gpu_single():
    ok, outs = run_ranks("_hpl_worker.py", 1, extra=["--n", "2048", "--nb", "256"], gpu=True, timeout=280)
    assert ok, "\n".join(outs)


@pytest.mark.gpu
@pytest.mark.multigpu
def test_hpl_mxp_gpu_multi():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(4, torch.cuda.device_count())
    ok, outs = run_ranks("_hpl_worker.py", world, extra=["--n", "4096", "--nb", "256"], gpu=True, timeout=280)
    assert ok, "\n".join(outs)
