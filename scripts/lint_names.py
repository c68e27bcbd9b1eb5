"""Tiny undefined-name check (no pyflakes in the image): names loaded in a function that are not
assigned anywhere in the module/function, imported, builtins or parameters."""
import ast
import builtins
import sys


def check(path):
    tree = ast.parse(open(path).read())
    module_names = set(dir(builtins))
    for node in ast.walk(tree):
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            for a in node.names:
                module_names.add((a.asname or a.name).split(".")[0])
        elif isinstance(node, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
            module_names.add(node.name)
            if not isinstance(node, ast.ClassDef):
                for a in node.args.args + node.args.kwonlyargs + node.args.posonlyargs:
                    module_names.add(a.arg)
                if node.args.vararg:
                    module_names.add(node.args.vararg.arg)
                if node.args.kwarg:
                    module_names.add(node.args.kwarg.arg)
        elif isinstance(node, ast.Name) and isinstance(node.ctx, (ast.Store, ast.Del)):
            module_names.add(node.id)
        elif isinstance(node, ast.ExceptHandler) and node.name:
            module_names.add(node.name)
        elif isinstance(node, ast.arg):
            module_names.add(node.arg)
    bad = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Name) and isinstance(node.ctx, ast.Load) and node.id not in module_names:
            bad.append((node.lineno, node.id))
    return bad


if __name__ == "__main__":
    rc = 0
    for p in sys.argv[1:]:
        for line, name in check(p):
            print("%s:%d: undefined name %s" % (p, line, name))
            rc = 1
    sys.exit(rc)
