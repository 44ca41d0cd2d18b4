"""HPCG retarget: CPU (stub transport) correctness at world 1 and 2; GPU variants in test_gpu_coll-style."""
import pytest
import torch

from _mp import run_ranks


@pytest.mark.parametrize("world", [1, 2])
def test_hpcg_cpu_stub(world):
    ok, outs = run_ranks("_hpcg_worker.py", world, extra=["--n", "8"], timeout=280)
    assert ok, "\n".join(outs)


@pytest.mark.gpu
def test_hpcg_gpu_single():
    ok, outs = run_ranks("_hpcg_worker.py", 1, extra=["--n", "32"], gpu=True, timeout=280)
    assert ok, "\n".join(outs)


@pytest.mark.gpu
@pytest.mark.multigpu
def test_hpcg_gpu_multi():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(4, torch.cuda.device_count())
    ok, outs = run_ranks("_hpcg_worker.py", world, extra=["--n", "32"], gpu=True, timeout=280)
    assert ok, "\n".join(outs)
