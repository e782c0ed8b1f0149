"""TEST INFRASTRUCTURE: run an UNMODIFIED Python 2 script of the reference under Python 3 with Python 2 semantics emulated.

No Python 2 interpreter exists offline, and /root/reference/bin/merge_cnts.py (the null-model roll-up)
only computes what its authors saw under Python 2: `a / b` on ints floors (:169, :256, :267 -- `it = it2/2` indexes a list),
numbers order before strings in `<`/`>` (:174, :263), and its output order is the iteration order of a Python 2 dict
(:207).  This harness parses the script's source as it lies in the reference tree, rewrites exactly those three constructs
in the AST -- Div -> py2 division, ordering comparisons -> py2 mixed-type ordering, `{}` / dict() -> a dict that iterates like
CPython 2.7's (open addressing, same growth rule and probe sequence, int/str hashes of a 64-bit build) -- and executes it
with the caller's argv.  Nothing else of the script changes; it is never copied into this repository.

Only tests/ and tests/golden/make_golden_rollup.py use this module.
"""
import ast
import sys


def py2_div(a, b):
    if isinstance(a, int) and isinstance(b, int) and not isinstance(a, bool) and not isinstance(b, bool):
        return a // b
    return a / b


def _rank(v):
    # CPython 2 default ordering between different types: None < numbers < everything else by type name
    if v is None:
        return (0, "")
    if isinstance(v, (int, float)):
        return (1, "")
    return (2, "str" if isinstance(v, str) else type(v).__name__)       # unicode_literals: text is 'unicode' > 'str' > 'list' ...


def py2_cmp(op, a, b):
    ra, rb = _rank(a), _rank(b)
    if ra == rb:
        x, y = a, b
    else:
        x, y = ra, rb
    return {"Lt": x < y, "Gt": x > y, "LtE": x <= y, "GtE": x >= y}[op]


def py2_hash(k):
    if isinstance(k, int):
        return -2 if k == -1 else ((k + (1 << 63)) % (1 << 64)) - (1 << 63)
    raise TypeError("py2 dict emulation: only int keys are iterated by the scripts this harness runs")


class Py2Dict(dict):
    """A dict whose ITERATION order is that of CPython 2.7 for int keys (Objects/dictobject.c: 8 slots, grows when
    fill * 3 >= size * 2 to used * 4 (used * 2 above 50000) rounded up to a power of two, probe i = 5 i + 1 + perturb).
    No deletions occur in the scripts it serves (a deletion would leave a dummy; refused)."""

    def __init__(self, *a, **kw):
        super().__init__()
        self._size, self._slots, self._fill = 8, [None] * 8, 0
        for k, v in dict(*a, **kw).items():
            self[k] = v

    def _reinsert(self, k):
        mask = self._size - 1
        h = py2_hash(k) & ((1 << 64) - 1)
        i = h & mask
        perturb = h
        while self._slots[i & mask] is not None:
            i = ((i << 2) + i + perturb + 1) & ((1 << 64) - 1)
            perturb >>= 5
        self._slots[i & mask] = k
        self._fill += 1

    def __setitem__(self, k, v):
        if k not in self:
            self._insert_fresh(k)
        super().__setitem__(k, v)

    def _insert_fresh(self, k):
        if not isinstance(k, int) or self._slots is None:
            self._slots = None
            return
        self._reinsert(k)
        if self._fill * 3 >= self._size * 2:
            used = self._fill
            want = used * (2 if used > 50000 else 4)
            new = 8
            while new <= want:
                new <<= 1
            old = [s for s in self._slots if s is not None]
            self._size, self._slots, self._fill = new, [None] * new, 0
            for kk in old:
                self._reinsert(kk)

    def setdefault(self, k, v=None):
        if k not in self:
            self[k] = v
        return super().__getitem__(k)

    def __delitem__(self, k):
        raise NotImplementedError("py2 dict emulation: deletions are not modelled")

    def pop(self, *a):
        raise NotImplementedError("py2 dict emulation: deletions are not modelled")

    def _order(self):
        if self._slots is None:
            return list(super().keys())
        return [s for s in self._slots if s is not None]

    def __iter__(self):
        return iter(self._order())

    def keys(self):
        return self._order()

    def items(self):
        return [(k, super(Py2Dict, self).__getitem__(k)) for k in self._order()]

    def values(self):
        return [super(Py2Dict, self).__getitem__(k) for k in self._order()]


class _Py2(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(ast.Call(func=ast.Name(id="__py2_div", ctx=ast.Load()), args=[node.left, node.right], keywords=[]), node)
        return node

    def visit_Compare(self, node):
        self.generic_visit(node)
        if len(node.ops) == 1 and type(node.ops[0]).__name__ in ("Lt", "Gt", "LtE", "GtE"):
            return ast.copy_location(ast.Call(func=ast.Name(id="__py2_cmp", ctx=ast.Load()),
                                              args=[ast.Constant(type(node.ops[0]).__name__), node.left, node.comparators[0]], keywords=[]), node)
        return node

    def visit_Dict(self, node):
        self.generic_visit(node)
        return ast.copy_location(ast.Call(func=ast.Name(id="__py2_dict", ctx=ast.Load()), args=[node], keywords=[]), node)


def run_script(path, argv):
    """Execute the Python 2 script at `path` with sys.argv = [path] + argv.  Returns its exit status (0 when it runs off the end)."""
    src = open(path).read()
    tree = ast.fix_missing_locations(_Py2().visit(ast.parse(src, path)))
    glb = {"__name__": "__main__", "__file__": path, "__py2_div": py2_div, "__py2_cmp": py2_cmp, "__py2_dict": Py2Dict}
    old = sys.argv
    sys.argv = [path] + list(argv)
    try:
        exec(compile(tree, path, "exec"), glb)
    except SystemExit as e:
        return int(e.code or 0)
    finally:
        sys.argv = old
        for v in list(glb.values()):              # the scripts leave their output files to interpreter exit
            if hasattr(v, "write") and hasattr(v, "close") and v not in (sys.stdout, sys.stderr):
                v.close()
    return 0


if __name__ == "__main__":
    sys.exit(run_script(sys.argv[1], sys.argv[2:]))
