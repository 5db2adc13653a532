"""Tiny undefined-name check (no pyflakes in this image): every Name that is loaded must be bound
somewhere in an enclosing function / module scope, a builtin, or a comprehension variable."""
import ast, builtins, sys

def check(path):
    tree = ast.parse(open(path).read(), path)
    problems = []
    class V(ast.NodeVisitor):
        def __init__(self): self.scopes = [set(dir(builtins))]
        def bind_targets(self, node):
            for n in ast.walk(node):
                if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)): self.scopes[-1].add(n.id)
        def collect(self, body_nodes, scope):
            for node in body_nodes:
                for n in ast.walk(node):
                    if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)): scope.add(n.name)
                    elif isinstance(n, ast.Import):
                        for a in n.names: scope.add((a.asname or a.name).split(".")[0])
                    elif isinstance(n, ast.ImportFrom):
                        for a in n.names: scope.add(a.asname or a.name)
                    elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)): scope.add(n.id)
                    elif isinstance(n, ast.arg): scope.add(n.arg)
                    elif isinstance(n, ast.ExceptHandler) and n.name: scope.add(n.name)
                    elif isinstance(n, (ast.Global, ast.Nonlocal)): scope.update(n.names)
        def visit_Module(self, node):
            s = set(); self.collect(node.body, s); self.scopes.append(s); self.generic_visit(node); self.scopes.pop()
        def visit_FunctionDef(self, node):
            s = set(); self.collect([node], s); self.scopes.append(s); self.generic_visit(node); self.scopes.pop()
        visit_AsyncFunctionDef = visit_FunctionDef
        visit_Lambda = visit_FunctionDef
        def visit_ClassDef(self, node):
            s = set(); self.collect(node.body, s); self.scopes.append(s); self.generic_visit(node); self.scopes.pop()
        def visit_Name(self, node):
            if isinstance(node.ctx, ast.Load) and not any(node.id in s for s in self.scopes):
                problems.append((node.lineno, node.id))
    V().visit(tree)
    return problems

bad = 0
for p in sys.argv[1:]:
    for ln, name in check(p):
        print(f"{p}:{ln}: undefined name {name}"); bad += 1
sys.exit(1 if bad else 0)
