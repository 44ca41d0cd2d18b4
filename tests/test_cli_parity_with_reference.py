"""Option-by-option comparison of the click command tree with the reference's `shipyard.py` (read as text, never imported).

Skipped when the reference checkout is not mounted (e.g. on the GPU box)."""
import os
import re

def _read(path):
    with open(path) as f:
        return f.read()


import click
import pytest

from batch_shipyard_b200 import cli

REF = "/root/reference/shipyard.py"
# options every command gets through shared decorators on both sides
COMMON = {"--configdir", "--credentials", "--config", "--pool", "--jobs", "--fs", "--monitor", "--federation", "--slurm", "--raw", "--yes",
          "--verbose", "--show-config"}


def _ours():
    out = {}

    def walk(g, path):
        for n, c in g.commands.items():
            if isinstance(c, click.Group):
                walk(c, path + [n])
            else:
                out[tuple(path + [n])] = {o for p in c.params if isinstance(p, click.Option) for o in p.opts if o.startswith("--")}
    walk(cli.cli, [])
    return out


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not mounted")
def test_every_reference_option_exists_on_the_same_command():
    src = _read(REF)
    blocks = re.findall(r"@(\w+)\.command\('([\w-]+)'\)(.*?)\ndef (\w+)\(", src, flags=re.S)
    assert len(blocks) == 105
    ours = _ours()
    assert len(ours) == 105
    missing = []
    for grp, name, decos, fn in blocks:
        opts = set(re.findall(r"@click\.option\(\s*'(--[\w-]+)'", decos)) - COMMON
        # the function name spells the command path: fed_jobs_del -> ("fed", "jobs", "del"); names keep their dashes
        # (sub-group commands drop the top group: nodes_count is `pool nodes count`, sas_create is `storage sas create`)
        joined = {k: "_".join(k).replace("-", "_") for k in ours}
        cands = ([k for k, j in joined.items() if j == fn] or [k for k, j in joined.items() if j.endswith("_" + fn) and k[-2] == grp]
                 or [k for k in ours if k[-1] == name and k[-2] == grp])                     # e.g. misc_mirror is `misc mirror-images`
        assert len(cands) == 1, (fn, cands)
        lost = sorted(o for o in opts if o not in ours[cands[0]])
        if lost:
            missing.append((" ".join(cands[0]), lost))
    assert missing == [], missing


@pytest.mark.skipif(not os.path.exists("/root/reference/docs/20-batch-shipyard-usage.md"), reason="reference checkout not mounted")
def test_raw_capable_commands_and_environment_variables_of_the_usage_guide_exist():
    """docs/20-batch-shipyard-usage.md lists the ~30 commands that support --raw and the SHIPYARD_* variables of the CLI."""
    txt = _read("/root/reference/docs/20-batch-shipyard-usage.md")
    block = txt[txt.index("The following commands support this option:"):txt.index("`--show-config` will output")]
    listed = [tuple(m.split()) for m in re.findall(r"\* `([a-z \-]+)`", block)]
    assert len(listed) >= 30
    ours = _ours()
    assert [c for c in listed if c not in ours] == []
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "batch_shipyard_b200", "cli.py")).read()
    env = set(re.findall(r"envvar='(SHIPYARD_[A-Z_]+)'", _read(REF)))
    assert len(env) >= 20 and sorted(e for e in env if e not in src) == []
