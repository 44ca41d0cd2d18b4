"""Static checks over the whole tree - the reference's only CI gate is `flake8 --select F,E,W` over its Python and `shellcheck` over
its shell scripts (/root/reference/.travis.yml:17-20, appveyor.yml:57-61).  Neither tool is in this image, so
the F-class checks that matter (syntax, unused imports, redefinitions, `except:` without a class, mutable default arguments) are
done with `ast`, shell scripts go through `bash -n`, and every YAML / JSON file in the tree must parse."""
import ast
import builtins
import json
import os
import subprocess

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SKIP_DIRS = {".git", "gpurun_out", "build", "baseline", "__pycache__", ".pytest_cache", ".hypothesis", "_native"}


def _files(*suffixes):
    out = []
    for dp, dn, fn in os.walk(ROOT):
        dn[:] = [d for d in dn if d not in SKIP_DIRS]
        out += [os.path.join(dp, f) for f in fn if f.endswith(suffixes)]
    return sorted(out)


PY = _files(".py") + [os.path.join(ROOT, "shipyard")]


def _tree(path):
    with open(path) as f:
        src = f.read()
    return src, ast.parse(src, path)


def test_python_sources_parse_and_have_no_unused_imports():
    bad = []
    for p in PY:
        src, tree = _tree(p)
        lines = src.splitlines()
        imported = {}
        for node in ast.walk(tree):
            if isinstance(node, ast.Import):
                for a in node.names:
                    imported[(a.asname or a.name).split(".")[0]] = node.lineno
            elif isinstance(node, ast.ImportFrom):
                for a in node.names:
                    if a.name != "*":
                        imported[a.asname or a.name] = node.lineno
        used = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name)}
        exported = set()
        for node in tree.body:                                      # names re-exported through __all__
            if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "__all__" for t in node.targets):
                exported = {e.value for e in ast.walk(node.value) if isinstance(e, ast.Constant) and isinstance(e.value, str)}
        for name, ln in imported.items():
            if name in used or name in exported or name == "annotations" or os.path.basename(p) == "__init__.py":
                continue
            if "noqa" in lines[ln - 1]:
                continue
            bad.append(f"{os.path.relpath(p, ROOT)}:{ln}: '{name}' imported but unused")
    assert bad == [], "\n".join(bad)


def test_no_redefinitions_bare_excepts_or_mutable_defaults():
    bad = []
    for p in PY:
        _, tree = _tree(p)
        rel = os.path.relpath(p, ROOT)
        for scope in [tree] + [n for n in ast.walk(tree) if isinstance(n, ast.ClassDef)]:
            seen = {}
            for node in scope.body:
                if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                    decorated = any(isinstance(d, ast.Attribute) and d.attr in ("setter", "getter", "deleter", "command", "group")
                                    or isinstance(d, ast.Call) for d in node.decorator_list)   # properties, click commands, overloads
                    if node.name in seen and not decorated:
                        bad.append(f"{rel}:{node.lineno}: redefinition of '{node.name}' from line {seen[node.name]}")
                    seen[node.name] = node.lineno
        for node in ast.walk(tree):
            if isinstance(node, ast.ExceptHandler) and node.type is None:
                bad.append(f"{rel}:{node.lineno}: bare 'except:'")
            if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
                for d in list(node.args.defaults) + [d for d in node.args.kw_defaults if d is not None]:
                    if isinstance(d, (ast.List, ast.Dict, ast.Set)):
                        bad.append(f"{rel}:{d.lineno}: mutable default argument")
    assert bad == [], "\n".join(bad)


def _bound_names(tree):
    b = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__builtins__", "__spec__", "__path__", "__class__"}
    for n in ast.walk(tree):
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            b.add(n.name)
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            a = n.args
            for x in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
                b.add(x.arg)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            b.add(n.id)
        elif isinstance(n, ast.Import):
            b.update((x.asname or x.name).split(".")[0] for x in n.names)
        elif isinstance(n, ast.ImportFrom):
            b.update(x.asname or x.name for x in n.names)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            b.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            b.update(n.names)
        elif isinstance(n, ast.MatchAs) and n.name:
            b.add(n.name)
    return b


def test_no_undefined_names():
    """flake8 F821, scope-insensitive: a name that is read somewhere in a module must be bound somewhere in that module (or be a
    builtin) - catches the NameError hiding in a branch no test walks through."""
    bad = []
    for p in PY:
        _, tree = _tree(p)
        if any(isinstance(n, ast.ImportFrom) and any(a.name == "*" for a in n.names) for n in ast.walk(tree)):
            continue
        bound = _bound_names(tree)
        bad += [f"{os.path.relpath(p, ROOT)}:{n.lineno}: undefined name '{n.id}'" for n in ast.walk(tree)
                if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in bound]
    assert bad == [], "\n".join(bad)


def test_the_checks_themselves_fire_on_a_bad_module():
    tree = ast.parse("import os\ndef f(a=[]):\n    try:\n        return undefined_thing\n    except:\n        pass\n")
    assert "undefined_thing" not in _bound_names(tree) and "os" in _bound_names(tree)
    assert any(isinstance(n, ast.ExceptHandler) and n.type is None for n in ast.walk(tree))
    assert any(isinstance(d, ast.List) for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) for d in n.args.defaults)


def test_shell_scripts_pass_bash_syntax_check():
    scripts = _files(".sh")
    assert scripts, "expected shell scripts in the tree (install.sh, bench/*.sh)"
    for p in scripts:
        r = subprocess.run(["bash", "-n", p], capture_output=True, text=True)
        assert r.returncode == 0, f"{os.path.relpath(p, ROOT)}: {r.stderr}"


def test_every_yaml_and_json_file_parses():
    n = 0
    for p in _files(".yaml", ".yml"):
        with open(p) as f:
            list(yaml.safe_load_all(f))
        n += 1
    for p in _files(".json"):
        with open(p) as f:
            txt = f.read().strip()
        if txt:
            for line in ([txt] if not p.endswith(".jsonl") else txt.splitlines()):
                json.loads(line)
        n += 1
    assert n > 150                                                   # recipes alone are 45 x 4 files


def test_python_sources_compile_to_bytecode():
    for p in PY:
        with open(p) as f:
            compile(f.read(), p, "exec")
