"""MPI launcher synthesis per runtime and autoscale scenario -> formula -> evaluation."""
import datetime

import pytest

from batch_shipyard_b200.config import settings as S
from batch_shipyard_b200.jobs import mpi as M
from batch_shipyard_b200.pool import autoscale as AS


def line(runtime, ppn, n=2, **kw):
    return M.construct_mpi_command(M.mpi_settings({"runtime": runtime, "processes_per_node": ppn, **{k: v for k, v in kw.items() if k in ("options", "executable_path")}}),
                                   n, "app", **{k: v for k, v in kw.items() if k not in ("options", "executable_path")})


def test_mpi_dialects():
    assert line("openmpi", 4)[0] == "mpirun --oversubscribe -host $AZ_BATCH_HOST_LIST -np 8 --map-by ppr:4:node --mca btl_tcp_if_include eth0 --allow-run-as-root app"
    assert line("mpich", 2)[0] == "mpirun -hosts $AZ_BATCH_HOST_LIST -np 4 -ppn 2 app"
    assert line("mvapich", 1)[0] == "mpirun -hosts $AZ_BATCH_HOST_LIST -np 2 -ppn 1 app"
    assert line("intelmpi", 3)[0] == "mpirun -hosts $AZ_BATCH_HOST_LIST -np 6 -perhost 3 app"
    cmd, _ = line("openmpi", "nvidia-smi -L | wc -l")
    assert "-np $(expr 2 \\* $(nvidia-smi -L | wc -l))" in cmd and "--map-by ppr:$(nvidia-smi -L | wc -l):node" in cmd
    cmd, env = line("intelmpi", 1, infiniband=True, rdma_class="sriov")
    assert env["I_MPI_FABRICS"] == "shm:ofi" and env["FI_PROVIDER"] == "mlx"
    _, env = line("intelmpi-ofa", 1, infiniband=True, rdma_class="sriov")
    assert env["I_MPI_FABRICS"] == "shm:ofa"
    _, env = line("intelmpi", 1, infiniband=True, rdma_class="networkdirect")
    assert env["I_MPI_FABRICS"] == "shm:dapl"
    cmd, _ = line("openmpi", 1, infiniband=True, rdma_class="sriov")
    assert "--mca pml ucx" in cmd and "UCX_NET_DEVICES=mlx5_0:1" in cmd
    cmd, env = line("openmpi", 8, n=1, infiniband=True, rdma_class="nvlink")
    assert "btl_tcp_if_include" not in cmd and env["SHIPYARD_COLL_TRANSPORT"] == "auto"
    cmd, _ = line("openmpi", 1, executable_path="/opt/MPI/bin/mpiexec", options=["-x FOO"])
    assert cmd.startswith("/opt/MPI/bin/mpiexec -x FOO --oversubscribe")       # case preserved (reference lower-cases: Q11)
    with pytest.raises(ValueError):
        M.mpi_settings({"runtime": "lam", "processes_per_node": 1})


def test_processes_per_node_resolution():
    assert M.resolve_processes_per_node("nvidia-smi -L | wc -l", 8) == 8
    assert M.resolve_processes_per_node(3, 8) == 3
    assert M.resolve_processes_per_node("echo 5", 0) == 5
    plan = M.make_launch_plan(2, 4, [0, 1, 2, 3, 4, 5, 6, 7], True, True)
    assert plan.world_size == 8 and plan.gpu_of_rank == list(range(8)) and plan.shim_face == "mpi"


def _pool(scn):
    return S.pool_settings({"pool_specification": {"id": "p", "vm_count": {"dedicated": 1, "low_priority": 0}, "max_tasks_per_node": 2,
                                                    "autoscale": {"evaluation_interval": "00:05:00", "scenario": scn}}})


def _metrics(active, n=30, now=1_000_000.0, period=30.0):
    m = AS.MetricsWindow(sample_period=period)
    for i in range(n):
        m.add("$ActiveTasks", now - (n - 1 - i) * period, active[i] if isinstance(active, list) else active)
        m.add("$PendingTasks", now - (n - 1 - i) * period, active[i] if isinstance(active, list) else active)
    m.current = {"$CurrentDedicatedNodes": 1, "$CurrentLowPriorityNodes": 0}
    return m, datetime.datetime.fromtimestamp(now)


def test_active_tasks_scenario_scales_with_load_and_increment_cap():
    scn = {"name": "active_tasks", "maximum_vm_count": {"dedicated": 8, "low_priority": 0},
           "maximum_vm_increment_per_evaluation": {"dedicated": 3, "low_priority": 0}, "bias_node_type": "dedicated"}
    pool = _pool(scn)
    text = AS.get_formula(pool)
    assert "$ActiveTasks" in text and "$TargetDedicatedNodes" in text and text.strip().endswith("$NodeDeallocationOption = taskcompletion;")
    m, now = _metrics(0)
    assert AS.FormulaInterpreter(m, now).run(text).target_dedicated == 1          # floor = vm_count
    m, now = _metrics(16)
    r = AS.FormulaInterpreter(m, now).run(text)
    assert r.target_dedicated == 4                                              # wants 8+, capped at current(1)+3
    m.current["$CurrentDedicatedNodes"] = 6
    assert AS.FormulaInterpreter(m, now).run(text).target_dedicated == 7          # 16 tasks / 2 per node - the floor node
    m2, now2 = _metrics(40)
    m2.current["$CurrentDedicatedNodes"] = 6
    assert AS.FormulaInterpreter(m2, now2).run(text).target_dedicated == 8        # ceiling


def test_calendar_scenarios_and_custom_formula():
    scn = {"name": "workday", "maximum_vm_count": {"dedicated": 8, "low_priority": 2}}
    text = AS.get_formula(_pool(scn))
    m, _ = _metrics(0)
    wed_10 = datetime.datetime(2026, 9, 23, 10, 0)     # a Wednesday
    sun_10 = datetime.datetime(2026, 9, 20, 10, 0)
    assert AS.FormulaInterpreter(m, wed_10).run(text).target_dedicated == 8
    assert AS.FormulaInterpreter(m, sun_10).run(text).target_dedicated == 1
    wk = AS.get_formula(_pool({"name": "weekend", "maximum_vm_count": {"dedicated": 8, "low_priority": 0}}))
    assert AS.FormulaInterpreter(m, sun_10).run(wk).target_dedicated == 8        # (always false in the reference's formula)
    custom = _pool({"name": "weekday", "maximum_vm_count": {"dedicated": 2, "low_priority": 0}})
    custom.autoscale.formula = "$TargetDedicatedNodes = min(4, max(1, avg($ActiveTasks.GetSample(TimeInterval_Minute * 5)) / 2));"
    assert AS.get_formula(custom) == custom.autoscale.formula                   # formula wins over scenario
    m, now = _metrics(6)
    assert AS.FormulaInterpreter(m, now).run(custom.autoscale.formula).target_dedicated == 3
    with pytest.raises(AS.FormulaError):
        AS.FormulaInterpreter(m, now).run("x = undefined_thing + 1;")
    with pytest.raises(ValueError):
        AS.get_formula(_pool({"name": "active_tasks", "maximum_vm_count": {"dedicated": 0, "low_priority": 0}}))


def test_formula_interpreter_fuzz_only_raises_formula_errors():
    """20 000 random token soups: the interpreter answers or raises FormulaError (a ValueError) — no TypeError / IndexError / hang."""
    import random
    import time
    from batch_shipyard_b200.pool.autoscale import FormulaError, FormulaInterpreter, MetricsWindow
    rng = random.Random(3)
    toks = ["$ActiveTasks", "$PendingTasks", "$CurrentDedicatedNodes", "$TargetDedicatedNodes", "$TargetLowPriorityNodes", "$NodeDeallocationOption",
            "x", "y", "=", "+", "-", "*", "/", "(", ")", "?", ":", "<", ">", "<=", ">=", "==", "!=", "&&", "||", "!", ",", ".", "GetSample",
            "GetSamplePercent", "(1)", "TimeInterval_Minute", "* 5", "min", "max", "avg", "val", "time", "()", "1", "0", "3.5", "taskcompletion",
            "t", ".hour", ".weekday", ";", "1e309", "--"]
    m = MetricsWindow()
    now = time.time()
    for i in range(20):
        m.add("$ActiveTasks", now - 60 * i, i); m.add("$PendingTasks", now - 60 * i, 2 * i)
    m.current.update({"$CurrentDedicatedNodes": 2, "$CurrentLowPriorityNodes": 0})
    answered = 0
    for _ in range(20000):
        f = " ".join(rng.choice(toks) for _ in range(rng.randint(1, 14)))
        if rng.random() < 0.5:
            f = "$TargetDedicatedNodes = " + f
        try:
            FormulaInterpreter(m).run(f)
            answered += 1
        except (FormulaError, ValueError):
            pass
    assert answered > 0
