#!/usr/bin/env python
"""Poor man's pyflakes (none in this image): report names that are read somewhere in a file but never bound anywhere
in it (assignment, import, def, argument, loop / with / except / comprehension target) and are not builtins.
Catches typos in code paths the CPU tests cannot execute (the GPU arm of bench.py).   python tools/lint_names.py FILES"""
import ast
import builtins
import sys


def check(path):
    tree = ast.parse(open(path).read(), path)
    bound, loads = set(dir(builtins)) | {"__file__", "__name__", "__path__"}, []
    for node in ast.walk(tree):
        if isinstance(node, ast.Name):
            if isinstance(node.ctx, ast.Load):
                loads.append(node)
            else:
                bound.add(node.id)
        elif isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            bound.add(node.name)
        elif isinstance(node, ast.arg):
            bound.add(node.arg)
        elif isinstance(node, (ast.Import, ast.ImportFrom)):
            for a in node.names:
                bound.add((a.asname or a.name).split(".")[0])
        elif isinstance(node, ast.ExceptHandler) and node.name:
            bound.add(node.name)
        elif isinstance(node, (ast.Global, ast.Nonlocal)):
            bound.update(node.names)
    bad = sorted({(n.id, n.lineno) for n in loads if n.id not in bound})
    for name, line in bad:
        print("%s:%d: name %r is never bound in this file" % (path, line, name))
    return len(bad)


if __name__ == "__main__":
    sys.exit(1 if sum(check(p) for p in sys.argv[1:]) else 0)
