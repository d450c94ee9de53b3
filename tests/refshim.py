"""TEST INFRASTRUCTURE: load the reference's pure-Python modules (Python-2 code)
under Python 3 WITHOUT copying them into the repository.

Only usable where /root/reference exists (the build container); tests that need
it skip elsewhere, and nothing under ``-m gpu`` uses it.  The module sources are
read from /root/reference/pylogeny/*.py, three textual py2->py3 patches are applied
in memory (SURVEY.md section 4 "shim recipe"), and the result is exec'd into module
objects registered under the bare names the reference imports
(`import newick, tree, base`, rearrangement.py:15).

The caller supplies the modules the reference expects to find beside it:
``scoring`` (ours), ``parsimony`` (ours or the reference's), ``alignment`` /
``executable`` stubs.
"""
import builtins
import collections
import collections.abc
import os
import sys
import types

REF = "/root/reference/pylogeny"


def available():
    return os.path.isdir(REF)


class _Nx1Graph(__import__("networkx").Graph):
    """networkx 1.x surface used by landscape.py:58-138,178,478,989 on a networkx 3 Graph."""

    @property
    def node(self):
        return self._node

    def edges_iter(self):
        return iter(self.edges())

    def neighbors(self, n):
        return list(super().neighbors(n))


_PATCHES = {
    "heuristic.py": [("except Exception,e:", "except Exception as e:")],
    "landscape.py": [("self.graph = networkx.Graph()", "self.graph = _Nx1Graph()")],
}


def _load(name, extra_patches=()):
    path = os.path.join(REF, name + ".py")
    with open(path) as fh:
        src = fh.read()
    for a, b in list(_PATCHES.get(name + ".py", [])) + list(extra_patches):
        src = src.replace(a, b)
    mod = types.ModuleType(name)
    mod.__file__ = path
    mod._Nx1Graph = _Nx1Graph
    sys.modules[name] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def install(provided, want=("base", "newick", "rearrangement", "tree")):
    """Install shimmed reference modules into sys.modules under their bare names.

    provided: dict name -> module object to register first (e.g. our `scoring`).
    Returns dict of all loaded reference modules. Call `uninstall()` afterwards."""
    builtins.xrange = range
    for n in ("Sized", "Iterable", "Container"):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))
    for k, v in provided.items():
        sys.modules[k] = v
    out = {}
    out["base"] = _load("base")
    if any(n in want for n in ("newick", "rearrangement", "tree")):
        _load_cycle(out)
    for n in ("parsimony", "landscape", "heuristic"):
        if n in want:
            out[n] = _load(n)
    return out


def _load_cycle(out):
    """tree -> imports newick -> imports rearrangement -> imports newick(partial), tree(partial)."""
    srcs = {}
    for n in ("tree", "newick", "rearrangement"):
        with open(os.path.join(REF, n + ".py")) as fh:
            srcs[n] = fh.read()
        m = types.ModuleType(n)
        m.__file__ = os.path.join(REF, n + ".py")
        out[n] = m

    # Emulate Python's import protocol for the cycle: a module is registered in
    # sys.modules before its body runs; `import x` of a partially initialised
    # module returns the partial object; `from x import y` needs y to exist.
    # newick.py:8 does `from rearrangement import topology`, so rearrangement's
    # body must have run before newick's -- achieved by importing tree first
    # (tree.py:11 -> newick -> rearrangement -> newick(partial), tree(partial)).
    class _Finder:
        def find_spec(self, fullname, path=None, target=None):
            if fullname in srcs and fullname not in sys.modules:
                import importlib.machinery
                import importlib.util
                return importlib.util.spec_from_loader(fullname, _Loader(fullname))
            return None

    class _Loader:
        def __init__(self, n):
            self.n = n

        def create_module(self, spec):
            return out[self.n]

        def exec_module(self, module):
            exec(compile(srcs[self.n], module.__file__, "exec"), module.__dict__)

    finder = _Finder()
    sys.meta_path.insert(0, finder)
    try:
        import importlib
        importlib.import_module("tree")
    finally:
        sys.meta_path.remove(finder)


def uninstall(names=("base", "newick", "rearrangement", "tree", "parsimony", "landscape", "heuristic", "scoring",
                     "alignment", "executable", "fitch", "pll", "libpllWrapper")):
    for n in names:
        sys.modules.pop(n, None)
