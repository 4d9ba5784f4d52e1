"""Cheap static checks of the Python plumbing that only runs on the GPU box (bench.py, the ctypes binding, the host
driver, the driver entry points): no undefined names, no local function sharing a name with a local variable."""
import ast
import builtins
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["bench.py", "__graft_entry__.py", "spmv-fpga_b200/spmvb.py", "spmv-fpga_b200/host_driver.py"]


def _module_names(tree):
    names = set(dir(builtins))
    for n in tree.body:
        if isinstance(n, (ast.Import, ast.ImportFrom)):
            names.update((a.asname or a.name).split(".")[0] for a in n.names)
        elif isinstance(n, (ast.FunctionDef, ast.ClassDef)):
            names.add(n.name)
        else:
            names.update(x.id for x in ast.walk(n) if isinstance(x, ast.Name) and isinstance(x.ctx, ast.Store))
    return names


def _functions(tree):
    for n in ast.walk(tree):
        if isinstance(n, ast.FunctionDef):
            yield n


@pytest.mark.parametrize("path", FILES)
def test_no_undefined_or_shadowed_names(path):
    tree = ast.parse(open(os.path.join(ROOT, path)).read())
    mod = _module_names(tree)
    class_names = {m.name for c in ast.walk(tree) if isinstance(c, ast.ClassDef) for m in c.body if isinstance(m, ast.FunctionDef)}
    problems = []
    for fn in (n for n in tree.body if isinstance(n, ast.FunctionDef)):
        local = set()
        for x in ast.walk(fn):
            if isinstance(x, ast.Name) and isinstance(x.ctx, (ast.Store, ast.Del)):
                local.add(x.id)
            elif isinstance(x, (ast.FunctionDef, ast.ClassDef)):
                local.add(x.name)
            elif isinstance(x, (ast.Import, ast.ImportFrom)):
                local.update((a.asname or a.name).split(".")[0] for a in x.names)
            elif isinstance(x, ast.ExceptHandler) and x.name:
                local.add(x.name)
            elif isinstance(x, ast.arg):
                local.add(x.arg)
        for x in ast.walk(fn):
            if isinstance(x, ast.Name) and isinstance(x.ctx, ast.Load) and x.id not in local and x.id not in mod:
                problems.append("%s:%d undefined name %s in %s()" % (path, x.lineno, x.id, fn.name))
    for fn in _functions(tree):
        nested = {n.name for n in fn.body if isinstance(n, ast.FunctionDef)}
        stored = {x.id for x in ast.walk(fn) if isinstance(x, ast.Name) and isinstance(x.ctx, ast.Store)}
        for name in nested & stored:
            problems.append("%s: %s() has a local function and a variable both called %s" % (path, fn.name, name))
    assert problems == [], "\n".join(problems)
    assert class_names is not None
