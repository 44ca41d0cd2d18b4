"""Task factories against the documented behaviour (reference docs/35-*.md golden cases)."""
import sys
import types


from batch_shipyard_b200.jobs import task_factory as tf


def cmds(task, config=None):
    return [t["command"] for t in tf.generate_tasks(config or {}, task)]


def test_product_single_and_nested():
    t = {"task_factory": {"parametric_sweep": {"product": [{"start": 0, "stop": 10, "step": 1}]}}, "command": '/bin/bash -c "sleep {0}"'}
    assert cmds(t) == [f'/bin/bash -c "sleep {i}"' for i in range(10)]
    t = {"task_factory": {"parametric_sweep": {"product": [{"start": 0, "stop": 3, "step": 1}, {"start": 100, "stop": 97, "step": -1}]}},
         "command": '/bin/bash -c "sleep {0}; sleep {1}"'}
    out = cmds(t)
    assert len(out) == 9 and out[0].endswith('sleep 0; sleep 100"') and out[2].endswith('sleep 0; sleep 98"') and out[8].endswith('sleep 2; sleep 98"')


def test_product_iterables_combinations_permutations_zip():
    t = {"task_factory": {"parametric_sweep": {"product_iterables": [["abc", "def", "ghi"], ["1", "2", "3"]]}}, "command": "echo {0}; sleep {1}"}
    out = cmds(t)
    assert len(out) == 9 and out[0] == "echo abc; sleep 1" and out[3] == "echo def; sleep 1" and out[8] == "echo ghi; sleep 3"
    t = {"task_factory": {"parametric_sweep": {"combinations": {"iterable": "abc", "length": 2, "replacement": False}}}, "command": "{0}{1}"}
    assert cmds(t) == ["ab", "ac", "bc"]
    t["task_factory"]["parametric_sweep"]["combinations"]["replacement"] = True
    assert cmds(t) == ["aa", "ab", "ac", "bb", "bc", "cc"]
    t = {"task_factory": {"parametric_sweep": {"permutations": {"iterable": "abc", "length": 2}}}, "command": "{0}{1}"}
    assert cmds(t) == ["ab", "ac", "ba", "bc", "ca", "cb"]
    t = {"task_factory": {"parametric_sweep": {"zip": ["ab", "01"]}}, "command": "{0}{1}"}
    assert cmds(t) == ["a0", "b1"]


def test_random_repeat():
    t = {"task_factory": {"random": {"generate": 3, "seed": 7, "integer": {"start": 0, "stop": 10, "step": 1}}}, "command": "sleep {}"}
    a, b = cmds(t), cmds(t)
    assert a == b and len(a) == 3 and all(0 <= int(c.split()[1]) < 10 for c in a)
    t = {"task_factory": {"random": {"generate": 4, "seed": 1, "distribution": {"uniform": {"a": 0.0, "b": 1.0}}}}, "command": "{}"}
    assert all(0.0 <= float(c) <= 1.0 for c in cmds(t))
    t = {"task_factory": {"repeat": 3}, "command": "sleep 1", "docker_image": "x"}
    out = list(tf.generate_tasks({}, t))
    assert len(out) == 3 and all("task_factory" not in o and o["command"] == "sleep 1" for o in out)


def test_custom_generator():
    m = types.ModuleType("sy_custom_gen")

    def generate(*args, **kwargs):
        for a in args:
            for x in range(0, int(a)):
                yield (x,)
    m.generate = generate
    sys.modules["sy_custom_gen"] = m
    t = {"task_factory": {"custom": {"module": "sy_custom_gen", "input_args": ["1", "2", "3"]}}, "command": '/bin/bash -c "sleep {}"'}
    assert [c[-2] for c in cmds(t)] == ["0", "0", "1", "0", "1", "2"]


def test_file_factory(tmp_path):
    root = tmp_path / "acct" / "mycontainer"
    (root / "archived").mkdir(parents=True)
    for n in ("test0.bin", "test1.bin", "archived/old0.bin", "archived/old1.bin", "skip.tmp"):
        (root / n).write_text("x")
    config = {"credentials": {"storage": {"mystorageaccount": {"local_path": str(tmp_path / "acct")}}}}
    t = {"task_factory": {"file": {"azure_storage": {"storage_account_settings": "mystorageaccount", "remote_path": "mycontainer", "exclude": ["*.tmp"]},
                                   "task_filepath": "file_path"}},
         "command": "echo full_path={file_path_with_container} file_path={file_path} file_name={file_name} noext={file_name_no_extension}"}
    out = list(tf.generate_tasks(config, t))
    assert len(out) == 4
    c = sorted(o["command"] for o in out)
    assert "full_path=mycontainer/archived/old0.bin file_path=archived/old0.bin file_name=old0.bin noext=old0" in c[0]
    assert all(o["resource_files"][-1]["blob_source"].startswith("file://") for o in out)
    assert sorted(o["resource_files"][-1]["file_path"] for o in out) == ["archived/old0.bin", "archived/old1.bin", "test0.bin", "test1.bin"]
