"""Pool autoscale: scenario -> formula text -> local evaluation.

The reference only *emits* Azure autoscale-formula text and lets the Batch
service evaluate it (/root/reference/convoy/autoscale.py:57-371).  There is no
service here, so this module has both halves: ``generate_formula`` turns a
scenario (active_tasks, pending_tasks, workday,
workday_with_offpeak_max_low_priority, weekday, weekend; bias / rebalance /
increment limits) into formula text, and ``FormulaInterpreter`` evaluates that
text — or a user's custom ``formula`` passthrough — against locally sampled
metrics to pick the number of active GPU "nodes".  One evaluation path for
generated and hand-written formulas.

Formula language (subset of the service's): ``name = expr;`` statements,
``+ - * / < <= > >= == != && || !``, ``c ? a : b``, functions ``min max avg val
time``, sample accessors ``$Metric.GetSample(n | interval [, pct])``,
``$Metric.GetSamplePercent(interval)``, ``TimeInterval_Second|Minute|Hour``,
``t.hour`` / ``t.weekday`` (0 = Sunday), outputs ``$TargetDedicatedNodes``,
``$TargetLowPriorityNodes``, ``$NodeDeallocationOption``.
"""
from __future__ import annotations

import datetime
import re
from dataclasses import dataclass, field
from typing import Any, Optional

_UNBOUNDED = 1 << 24


# ---------------------------------------------------------------------------
# formula generation
# ---------------------------------------------------------------------------
@dataclass
class _Limits:
    max_tasks_per_node: int
    min_dedicated: int
    min_low_priority: int
    max_dedicated: int
    max_low_priority: int
    inc_dedicated: int
    inc_low_priority: int


def _limits(pool) -> _Limits:
    sc = pool.autoscale.scenario
    mx_d = _UNBOUNDED if sc.max_dedicated < 0 else sc.max_dedicated
    mx_l = _UNBOUNDED if sc.max_low_priority < 0 else sc.max_low_priority
    if mx_d < pool.vm_dedicated or mx_l < pool.vm_low_priority:
        raise ValueError("autoscale maximum_vm_count is below the pool's vm_count (the minimum)")
    inc_d = _UNBOUNDED if sc.inc_dedicated <= 0 else sc.inc_dedicated
    inc_l = _UNBOUNDED if sc.inc_low_priority <= 0 else sc.inc_low_priority
    return _Limits(pool.max_tasks_per_node, pool.vm_dedicated, pool.vm_low_priority, mx_d, mx_l, inc_d, inc_l)


def _task_scenario(pool, lm: _Limits) -> list[str]:
    sc = pool.autoscale.scenario
    metric = "$ActiveTasks" if sc.name == "active_tasks" else "$PendingTasks"
    look = int(sc.sample_lookback_interval.total_seconds())
    pct = sc.required_sample_percentage
    s = [f"window = TimeInterval_Second * {look}"]
    if sc.bias_last_sample:
        s += [f"havePct = {metric}.GetSamplePercent(window)",
              f"newest = val({metric}.GetSample(1), 0)",
              f"windowAvg = havePct < {pct} ? max(0, newest) : avg({metric}.GetSample(window))",
              # rising load follows the newest sample, falling load decays through the average
              f"load = havePct < {pct} ? max(0, newest) : (newest < windowAvg ? avg(newest, windowAvg) : max(newest, windowAvg))",
              "wantNodes = load / tasksPerNode"]
    else:
        s += [f"load = avg({metric}.GetSample(window, {pct}))",
              "wantNodes = load / tasksPerNode",
              "wantNodes = (load > 0 && wantNodes < 1) ? 1 : wantNodes"]
    if sc.rebalance_preemption_percentage is not None:
        if sc.bias_last_sample:
            s += ["prePct = $PreemptedNodeCount.GetSamplePercent(window)",
                  "preNewest = val($PreemptedNodeCount.GetSample(1), 0)",
                  "preAvg = avg($PreemptedNodeCount.GetSample(window))",
                  f"preempted = prePct < {pct} ? max(0, preNewest) : (preNewest > preAvg ? avg(preNewest, preAvg) : min(preNewest, preAvg))"]
        else:
            s += [f"preempted = avg($PreemptedNodeCount.GetSample(window, {pct}))"]
        s += ["haveNodes = $CurrentDedicatedNodes + $CurrentLowPriorityNodes",
              "preemptedShare = haveNodes > 0 ? preempted / haveNodes : 0",
              f"shift = preemptedShare >= {sc.rebalance_preemption_percentage}"]
    else:
        s += ["preempted = 0", "shift = 0 == 1"]
    s += ["capDedicated = min($CurrentDedicatedNodes + stepDedicated, ceilDedicated)",
          "capLowPri = min($CurrentLowPriorityNodes + stepLowPri, ceilLowPri)",
          "wantNodes = max(0, wantNodes - floorDedicated - floorLowPri)"]
    bias = sc.bias_node_type
    if bias == "auto":
        s += ["parts = (ceilDedicated == 0 || ceilLowPri == 0) ? 1 : 2",
              "ded = wantNodes / parts",
              "ded = (ded > 0 && ded < 1) ? 1 : ded",
              "ded = max(floorDedicated, min(ded, capDedicated))"]
    elif bias == "dedicated":
        s += ["ded = max(floorDedicated, min(wantNodes, capDedicated))"]
    elif bias == "low_priority":
        s += ["low = max(floorLowPri, min(wantNodes, capLowPri))",
              "rest = max(0, wantNodes - low)",
              "moved = (shift && rest > 0) ? min(preempted, rest) : 0",
              "low = min(low + moved, capLowPri)",
              "rest = max(0, wantNodes - low)",
              "ded = max(floorDedicated, min(rest, capDedicated))",
              "$TargetLowPriorityNodes = low", "$TargetDedicatedNodes = ded"]
        return s
    else:
        raise ValueError(f"bad autoscale bias_node_type '{bias}'")
    s += ["rest = max(0, wantNodes - ded)",
          "moved = (shift && rest > 0) ? min(preempted, rest) : 0",
          "ded = min(ded + moved, capDedicated)",
          "rest = max(0, wantNodes - ded)",
          "low = max(floorLowPri, min(rest, capLowPri))",
          "$TargetDedicatedNodes = ded", "$TargetLowPriorityNodes = low"]
    return s


def _calendar_scenario(pool) -> list[str]:
    sc = pool.autoscale.scenario
    s = ["t = time()"]
    work = "inHours = t.hour >= hourFrom && t.hour <= hourTo"
    week = "inWeek = t.weekday >= dayFrom && t.weekday <= dayTo"
    if sc.name in ("workday", "workday_with_offpeak_max_low_priority"):
        s += [work, week, "peak = inWeek && inHours"]
    elif sc.name == "weekday":
        s += ["peak = t.weekday >= dayFrom && t.weekday <= dayTo"]
    elif sc.name == "weekend":
        s += ["peak = t.weekday < dayFrom || t.weekday > dayTo"]
    else:
        raise ValueError(f"bad autoscale scenario '{sc.name}'")
    bias = sc.bias_node_type
    if bias not in ("auto", "dedicated", "low_priority"):
        raise ValueError(f"bad autoscale bias_node_type '{bias}'")
    if sc.name == "workday_with_offpeak_max_low_priority":
        s += ["$TargetLowPriorityNodes = ceilLowPri",
              "$TargetDedicatedNodes = floorDedicated" if bias == "low_priority"
              else "$TargetDedicatedNodes = peak ? ceilDedicated : floorDedicated"]
        return s
    ded_peak = "$TargetDedicatedNodes = peak ? ceilDedicated : floorDedicated"
    low_peak = "$TargetLowPriorityNodes = peak ? ceilLowPri : floorLowPri"
    if bias == "auto":
        s += [ded_peak, low_peak]
    elif bias == "dedicated":
        s += [ded_peak, "$TargetLowPriorityNodes = floorLowPri"]
    else:
        s += ["$TargetDedicatedNodes = floorDedicated", low_peak]
    return s


def generate_formula(pool) -> str:
    """Formula text for ``pool.autoscale.scenario`` (see module docstring for the language)."""
    sc = pool.autoscale.scenario
    lm = _limits(pool)
    head = [f"tasksPerNode = {lm.max_tasks_per_node}", f"floorDedicated = {lm.min_dedicated}",
            f"floorLowPri = {lm.min_low_priority}", f"ceilDedicated = {lm.max_dedicated}",
            f"ceilLowPri = {lm.max_low_priority}"]
    if sc.name in ("active_tasks", "pending_tasks"):
        head += [f"stepDedicated = {lm.inc_dedicated}", f"stepLowPri = {lm.inc_low_priority}"]
        body = _task_scenario(pool, lm)
    else:
        head += [f"dayFrom = {sc.weekday_start}", f"dayTo = {sc.weekday_end}", f"hourFrom = {sc.workhour_start}",
                 f"hourTo = {sc.workhour_end}"]
        body = _calendar_scenario(pool)
    tail = [f"$NodeDeallocationOption = {sc.node_deallocation_option}"]
    return ";\n".join(head + body + tail) + ";"


def get_formula(pool) -> str:
    """A custom ``formula`` wins over a ``scenario`` (same precedence as the reference)."""
    if pool.autoscale is None:
        raise ValueError("pool has no autoscale settings")
    if pool.autoscale.formula:
        return pool.autoscale.formula
    if pool.autoscale.scenario is None:
        raise ValueError("autoscale needs a scenario or a formula")
    return generate_formula(pool)


# ---------------------------------------------------------------------------
# metrics
# ---------------------------------------------------------------------------
@dataclass
class MetricsWindow:
    """Time-stamped samples per metric (newest last).  sample_period: nominal seconds between samples."""
    sample_period: float = 30.0
    samples: dict = field(default_factory=dict)   # name -> list[(ts: float, value: float)]
    current: dict = field(default_factory=dict)   # $CurrentDedicatedNodes etc.

    def add(self, name: str, ts: float, value: float) -> None:
        self.samples.setdefault(name, []).append((float(ts), float(value)))

    def window(self, name: str, now: float, seconds: float) -> list[float]:
        return [v for (t, v) in self.samples.get(name, []) if now - seconds <= t <= now]

    def last(self, name: str, n: int) -> list[float]:
        return [v for (_, v) in self.samples.get(name, [])[-n:]]

    def percent(self, name: str, now: float, seconds: float) -> float:
        expected = max(1.0, seconds / self.sample_period)
        return min(100.0, 100.0 * len(self.window(name, now, seconds)) / expected)


@dataclass
class AutoscaleResult:
    target_dedicated: int
    target_low_priority: int
    node_deallocation_option: str
    variables: dict


# ---------------------------------------------------------------------------
# interpreter
# ---------------------------------------------------------------------------
_TOKEN = re.compile(r"\s*(?:(\d+\.\d+|\d+)|(\$?[A-Za-z_][A-Za-z_0-9]*)|(&&|\|\||==|!=|<=|>=|[-+*/<>!?:(),.=;]))")


class FormulaError(ValueError):
    pass


class _Interval:
    def __init__(self, seconds: float):
        self.seconds = float(seconds)

    def __mul__(self, k):
        return _Interval(self.seconds * float(k))

    __rmul__ = __mul__


class FormulaInterpreter:
    def __init__(self, metrics: MetricsWindow, now: Optional[datetime.datetime] = None):
        self.m = metrics
        self.now = now or datetime.datetime.now()
        self.now_ts = self.now.timestamp()
        self.vars: dict[str, Any] = {"TimeInterval_Second": _Interval(1), "TimeInterval_Minute": _Interval(60),
                                     "TimeInterval_Hour": _Interval(3600), "TimeInterval_Zero": _Interval(0)}
        for k, v in metrics.current.items():
            self.vars[k] = float(v)

    # -- tokenizer / parser (precedence climbing) ----------------------------
    def _tokens(self, text: str) -> list[str]:
        out, pos = [], 0
        text = re.sub(r"//[^\n]*", "", text)
        while pos < len(text):
            if text[pos:].strip() == "":
                break
            mt = _TOKEN.match(text, pos)
            if not mt:
                raise FormulaError(f"cannot tokenize formula near: {text[pos:pos + 20]!r}")
            out.append(mt.group(1) or mt.group(2) or mt.group(3))
            pos = mt.end()
        return out

    def run(self, formula: str) -> AutoscaleResult:
        for stmt in [s.strip() for s in formula.split(";")]:
            if not stmt:
                continue
            self.toks, self.i = self._tokens(stmt), 0
            if len(self.toks) < 3 or self.toks[1] != "=":
                raise FormulaError(f"expected 'name = expression' in: {stmt!r}")
            name = self.toks[0]
            self.i = 2
            if name == "$NodeDeallocationOption":
                self.vars[name] = self.toks[2]
                continue
            try:
                val = self._ternary()
            except (TypeError, IndexError, AttributeError, ZeroDivisionError, OverflowError, RecursionError) as e:
                # a well-tokenised but ill-typed statement (an interval used as a number, a function without arguments, ...)
                raise FormulaError(f"cannot evaluate {stmt!r}: {e}") from e
            if self.i != len(self.toks):
                raise FormulaError(f"trailing tokens in: {stmt!r}")
            self.vars[name] = val
        d = self.vars.get("$TargetDedicatedNodes", self.m.current.get("$CurrentDedicatedNodes", 0))
        lp = self.vars.get("$TargetLowPriorityNodes", self.m.current.get("$CurrentLowPriorityNodes", 0))
        try:
            d, lp = float(d), float(lp)
            if d != d or lp != lp or abs(d) == float("inf") or abs(lp) == float("inf"):
                raise ValueError("not a finite number")
        except (TypeError, ValueError) as e:
            raise FormulaError(f"$TargetDedicatedNodes / $TargetLowPriorityNodes must evaluate to finite numbers: {e}") from e
        return AutoscaleResult(max(0, int(float(d))), max(0, int(float(lp))),
                               str(self.vars.get("$NodeDeallocationOption", "requeue")),
                               {k: v for k, v in self.vars.items() if not isinstance(v, (_Interval, list))})

    def _peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else None

    def _eat(self, t=None):
        tok = self._peek()
        if tok is None or (t is not None and tok != t):
            raise FormulaError(f"expected {t!r}, found {tok!r}")
        self.i += 1
        return tok

    def _ternary(self):
        c = self._binary(0)
        if self._peek() == "?":
            self._eat("?")
            a = self._ternary()
            self._eat(":")
            b = self._ternary()
            return a if self._truth(c) else b
        return c

    _PREC = {"||": 1, "&&": 2, "==": 3, "!=": 3, "<": 4, "<=": 4, ">": 4, ">=": 4, "+": 5, "-": 5, "*": 6, "/": 6}

    def _binary(self, minp):
        lhs = self._unary()
        while True:
            op = self._peek()
            p = self._PREC.get(op)
            if p is None or p < minp:
                return lhs
            self._eat()
            rhs = self._binary(p + 1)
            lhs = self._apply(op, lhs, rhs)

    @staticmethod
    def _truth(v) -> bool:
        return bool(v) and v != 0

    def _apply(self, op, a, b):
        if op == "||":
            return 1.0 if self._truth(a) or self._truth(b) else 0.0
        if op == "&&":
            return 1.0 if self._truth(a) and self._truth(b) else 0.0
        if op == "*" and (isinstance(a, _Interval) or isinstance(b, _Interval)):
            return a * b if isinstance(a, _Interval) else b * a
        a, b = float(a), float(b)
        if op == "+": return a + b
        if op == "-": return a - b
        if op == "*": return a * b
        if op == "/": return a / b if b != 0 else 0.0
        return 1.0 if {"==": a == b, "!=": a != b, "<": a < b, "<=": a <= b, ">": a > b, ">=": a >= b}[op] else 0.0

    def _unary(self):
        t = self._peek()
        if t == "!":
            self._eat(); return 0.0 if self._truth(self._unary()) else 1.0
        if t == "-":
            self._eat(); return -float(self._unary())
        return self._postfix(self._primary())

    def _args(self) -> list:
        self._eat("(")
        args = []
        if self._peek() != ")":
            args.append(self._ternary())
            while self._peek() == ",":
                self._eat(","); args.append(self._ternary())
        self._eat(")")
        return args

    def _primary(self):
        t = self._eat()
        if t == "(":
            v = self._ternary(); self._eat(")"); return v
        if re.match(r"^\d", t):
            return float(t)
        if self._peek() == "(" and not t.startswith("$"):
            return self._call(t, self._args())
        if t.startswith("$") and self._peek() == ".":
            return ("metric", t)
        if t in self.vars:
            return self.vars[t]
        raise FormulaError(f"unknown identifier '{t}'")

    def _postfix(self, v):
        while self._peek() == ".":
            self._eat(".")
            attr = self._eat()
            if isinstance(v, tuple) and v[0] == "metric":
                v = self._metric_call(v[1], attr, self._args())
            elif isinstance(v, datetime.datetime):
                if attr == "hour": v = float(v.hour)
                elif attr == "weekday": v = float((v.weekday() + 1) % 7)   # 0 = Sunday
                elif attr == "minute": v = float(v.minute)
                else: raise FormulaError(f"unknown time attribute '{attr}'")
            else:
                raise FormulaError(f"cannot take '.{attr}' of {v!r}")
        return v

    def _flat(self, args) -> list[float]:
        out: list[float] = []
        for a in args:
            out.extend(float(x) for x in a) if isinstance(a, list) else out.append(float(a))
        return out

    def _call(self, fn, args):
        if fn == "time":
            return self.now
        vals = self._flat(args)
        if fn == "min": return min(vals) if vals else 0.0
        if fn == "max": return max(vals) if vals else 0.0
        if fn == "avg": return sum(vals) / len(vals) if vals else 0.0
        if fn == "sum": return sum(vals)
        if fn == "val":
            vec, idx = args[0], int(float(args[1]))
            return float(vec[idx]) if isinstance(vec, list) and idx < len(vec) else 0.0
        if fn == "len": return float(len(args[0])) if isinstance(args[0], list) else 1.0
        raise FormulaError(f"unknown function '{fn}'")

    def _metric_call(self, metric, method, args):
        if method == "GetSamplePercent":
            iv = args[0]
            return self.m.percent(metric, self.now_ts, iv.seconds if isinstance(iv, _Interval) else float(iv))
        if method == "GetSample":
            a0 = args[0]
            if isinstance(a0, _Interval):
                w = self.m.window(metric, self.now_ts, a0.seconds)
                if len(args) > 1 and self.m.percent(metric, self.now_ts, a0.seconds) < float(args[1]):
                    # not enough samples for the required percentage: behave like an empty vector
                    return []
                return w
            return list(reversed(self.m.last(metric, int(float(a0)))))   # newest first
        raise FormulaError(f"unknown sample accessor '{method}'")


def evaluate(pool, metrics: MetricsWindow, now: Optional[datetime.datetime] = None) -> AutoscaleResult:
    """Evaluate the pool's formula/scenario against local metrics; clamp to what the box has."""
    res = FormulaInterpreter(metrics, now).run(get_formula(pool))
    return res
