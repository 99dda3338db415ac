"""bench.py, __graft_entry__.py and the tools the GPU round runs cannot be executed without a GPU; what CAN be checked on CPU is
that every name a function reads is bound somewhere (module level, the function itself, an enclosing function, or builtins) --
the class of mistake (a helper moved out of `main` that still uses one of main's local imports) that otherwise only shows up on
the GPU box, in the driver's run."""
import ast
import builtins
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["bench.py", "__graft_entry__.py", "tools/make_profiles.py", "tools/bench_summary.py", "tools/small_batch_probe.py",
         "tools/step_trace.py", "tools/step_trace_summary.py"]


def _bound_names(node):
    """Names bound directly in `node`'s own scope (not in nested functions / classes)."""
    out = set()

    def visit(n, top):
        for c in ast.iter_child_nodes(n):
            if isinstance(c, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                out.add(c.name)
                continue   # its body is another scope
            if isinstance(c, ast.Lambda):
                continue
            if isinstance(c, ast.Name) and isinstance(c.ctx, (ast.Store, ast.Del)):
                out.add(c.id)
            elif isinstance(c, (ast.Import, ast.ImportFrom)):
                for a in c.names:
                    out.add((a.asname or a.name).split(".")[0])
            elif isinstance(c, ast.ExceptHandler) and c.name:
                out.add(c.name)
            elif isinstance(c, (ast.Global, ast.Nonlocal)):
                out.update(c.names)
            visit(c, False)

    if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
        a = node.args
        for arg in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
            out.add(arg.arg)
    visit(node, True)
    return out


def _undefined(tree):
    missing = []

    def walk(node, scopes):
        here = scopes + [_bound_names(node)]
        for c in ast.iter_child_nodes(node):
            check(c, here)

    def check(n, scopes):
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            walk(n, scopes)
            return
        if isinstance(n, ast.ClassDef):
            walk(n, scopes)
            return
        if isinstance(n, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
            targets = set()
            for g in n.generators:
                for t in ast.walk(g.target):
                    if isinstance(t, ast.Name):
                        targets.add(t.id)
            for c in ast.iter_child_nodes(n):
                check(c, scopes + [targets])
            return
        if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load):
            if not any(n.id in s for s in scopes) and not hasattr(builtins, n.id) and n.id not in ("__file__", "__name__", "__doc__"):
                missing.append((n.id, n.lineno))
        for c in ast.iter_child_nodes(n):
            check(c, scopes)

    walk(tree, [])
    return missing


@pytest.mark.parametrize("rel", FILES)
def test_every_name_read_is_bound_somewhere(rel):
    path = os.path.join(ROOT, rel)
    tree = ast.parse(open(path).read(), filename=rel)
    missing = _undefined(tree)
    assert not missing, f"{rel}: names read but never bound: {sorted(set(missing))}"


def test_the_checker_sees_the_mistake_it_is_there_for():
    src = "def main():\n    import json as j\n    helper()\n\ndef helper():\n    return j.dumps({})\n"
    assert [m[0] for m in _undefined(ast.parse(src))] == ["j"]
