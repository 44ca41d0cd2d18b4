"""HPMLA retarget (SymSGD linear learner): composition math on the stub transport, the data shredder, the recipe end to end."""
import json
import os
import subprocess
import sys

import pytest

from _mp import run_ranks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RECIPE = os.path.join(ROOT, "recipes", "HPMLA-CPU-OpenMPI")


@pytest.mark.parametrize("world", [1, 2, 3])
def test_symsgd_composition_matches_sequential_sgd(world):
    ok, outs = run_ranks("_hpmla_worker.py", world, timeout=280)
    assert ok, "\n".join(outs)


def test_shred_and_train_from_files(tmp_path):
    prefix = str(tmp_path / "shards" / "train")
    p = subprocess.run([sys.executable, os.path.join(RECIPE, "shred_data.py"), "--synthetic", "1200", "--dim", "16", "--out-prefix", prefix,
                        "--node-count", "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert p.returncode == 0, p.stdout
    lines = [sum(1 for _ in open(f"{prefix}.{r}")) for r in range(2)]
    assert lines == [600, 600]
    env = dict(os.environ, SHIPYARD_GPU="-1")
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    logs = str(tmp_path / "models")
    p = subprocess.run([sys.executable, os.path.join(RECIPE, "supersgd.py"), "-l", "0.1", "-k", "8", "-m", "1e-2", "-e", "4", "-r", "3", "-f", prefix,
                        "-t", "1", "-g", "2", "-d", logs, "--dim", "16", "--batch", "32"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=200, env=env)
    assert p.returncode == 0, p.stdout
    out = json.loads(p.stdout.strip().splitlines()[-1])
    assert out["examples_per_rank"] == 600 and out["final_loss"] < out["first_epoch_loss"] and out["final_accuracy"] > 0.85
    assert sorted(os.listdir(logs)) == ["global_model_epoch_2.txt", "global_model_epoch_4.txt"]
    assert sum(1 for _ in open(os.path.join(logs, "global_model_epoch_4.txt"))) == 16
