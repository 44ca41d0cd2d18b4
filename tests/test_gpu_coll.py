"""GPU collectives: every kernel vs a PyTorch fp64 reference, 1 GPU and (when present) 2/4/8 GPUs."""
import pytest
import torch

from _mp import run_ranks

pytestmark = pytest.mark.gpu


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_single_gpu_collectives():
    ok, outs = run_ranks("_coll_worker.py", 1, extra=["--quick"], gpu=True, timeout=300)
    assert ok, "\n".join(o[-3000:] for o in outs)
    assert "transport=p2p" in outs[0] or "transport=nvls" in outs[0]


@pytest.mark.multigpu
@pytest.mark.parametrize("transport", ["auto", "p2p"])
def test_multi_gpu_collectives(transport):
    world = min(_ngpu(), 8)
    ok, outs = run_ranks("_coll_worker.py", world, extra=["--transport", transport], gpu=True, timeout=600)
    assert ok, "\n".join(o[-3000:] for o in outs)
