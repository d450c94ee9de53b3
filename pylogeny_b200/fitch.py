"""Drop-in for the reference's `fitch` extension module (fitch.cpp:274-346).

    calculateCost(newick, profiles, weights, taxa) -> int

Same argument meaning as fitch_cost(): a Newick string, a list of profile
strings (one char per alignment row), a list of integer weights, a dict
taxon label -> row index.  The work is done by the sm_100a Fitch kernel behind
the C-ABI (include/pylogeny_b200.h: plg_fitch_*); there is no CPU fallback.

Error behaviour: the reference returns NULL without setting an exception on a
length mismatch or wrongly-typed arguments (fitch.cpp:285,300,308,324), which
CPython surfaces as SystemError -- raised here as SystemError too.  An unknown
leaf label makes the reference abort through an uncaught std::out_of_range
(fitch.cpp:255); here it raises KeyError instead of killing the interpreter.
"""
import ctypes as C

import numpy as np

from . import _lib

__all__ = ["calculateCost", "FitchAlignment"]


class FitchAlignment(object):
    """Device-resident bit-plane packed pattern set (plg_fitch_aln)."""

    def __init__(self, profiles, weights):
        if len(profiles) != len(weights):
            raise SystemError("error return without exception set")  # fitch.cpp:285
        n = len(profiles)
        try:
            enc = [p.encode("latin-1") for p in profiles]
            w = np.asarray([int(x) for x in weights], dtype=np.int64)
        except (AttributeError, TypeError, ValueError):
            raise SystemError("error return without exception set")  # fitch.cpp:300,308
        arr = (C.c_char_p * max(n, 1))(*enc)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().plg_fitch_aln_create(n, arr, w.ctypes.data_as(C.POINTER(C.c_int64)),
                                                    C.byref(self._h)))
        self.n_profiles = n

    @classmethod
    def from_dense(cls, states, weights):
        """states: uint8 [n_patterns, n_rows] symbol matrix; weights: int64 [n_patterns]."""
        self = cls.__new__(cls)
        states = np.ascontiguousarray(states, dtype=np.uint8)
        w = np.ascontiguousarray(weights, dtype=np.int64)
        if states.ndim != 2 or states.shape[0] != w.shape[0]:
            raise SystemError("error return without exception set")
        self._h = C.c_void_p()
        _lib.check(_lib.lib().plg_fitch_aln_create_dense(states.shape[0], states.shape[1], states.ctypes.data,
                                                          w.ctypes.data, C.byref(self._h)))
        self.n_profiles = states.shape[0]
        return self

    def info(self):
        P, nc, k = C.c_int64(), C.c_int(), C.c_int()
        _lib.check(_lib.lib().plg_fitch_aln_info(self._h, C.byref(P), C.byref(nc), C.byref(k)))
        return {"n_prof": P.value, "n_cols": nc.value, "n_planes": k.value}

    def score_newicks(self, newicks, taxa):
        """Fitch cost of each Newick string (k-ary nodes allowed). Returns int32 array."""
        if not isinstance(taxa, dict):
            raise SystemError("error return without exception set")  # fitch.cpp:324
        names = list(taxa.keys())
        nt = len(names)
        narr = (C.c_char_p * max(nt, 1))(*[str(k).encode("latin-1") for k in names])
        idx = np.asarray([int(taxa[k]) for k in names], dtype=np.int32)
        n = len(newicks)
        sarr = (C.c_char_p * max(n, 1))(*[s.encode("latin-1") for s in newicks])
        out = np.zeros(max(n, 1), dtype=np.int32)
        rc = _lib.lib().plg_fitch_score_newicks(self._h, n, sarr, nt, narr,
                                                idx.ctypes.data_as(C.POINTER(C.c_int32)), out.ctypes.data, None)
        if rc == _lib.ERR_LABEL:
            raise KeyError(_lib.lib().plg_last_error().decode())
        _lib.check(rc)
        return out[:n]

    def score_trees(self, desc, tip_rows=None):
        """Batched descriptor path: desc int32 [n_trees, n_tips-1, 2] (see include/pylogeny_b200.h)."""
        desc = np.ascontiguousarray(desc, dtype=np.int32)
        if desc.ndim == 2:
            desc = desc[None]
        n_trees, ni, two = desc.shape
        assert two == 2
        rows = None
        if tip_rows is not None:
            rows = np.ascontiguousarray(tip_rows, dtype=np.int32)
        out = np.zeros(max(n_trees, 1), dtype=np.int32)
        _lib.check(_lib.lib().plg_fitch_score_trees(self._h, n_trees, ni + 1, desc.ctypes.data,
                                                     rows.ctypes.data if rows is not None else None,
                                                     out.ctypes.data, None))
        return out[:n_trees]

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().plg_fitch_aln_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_last = None  # (profiles list, weights list, FitchAlignment): reuse the device copy across calls


def calculateCost(newick, profiles, weights, taxa):
    """Given a Newick string tree, calculate the parsimony cost (fitch.cpp:338-342)."""
    global _last
    if not isinstance(newick, str) or not isinstance(profiles, list) or not isinstance(weights, list):
        raise TypeError("calculateCost(newick:str, profiles:list, weights:list, taxa:dict)")
    if len(profiles) != len(weights):
        raise SystemError("error return without exception set")
    if not isinstance(taxa, dict):
        raise SystemError("error return without exception set")
    if _last is None or _last[0] != profiles or _last[1] != weights:
        if _last is not None:
            _last[2].close()
        _last = (list(profiles), list(weights), FitchAlignment(profiles, weights))
    return int(_last[2].score_newicks([newick], taxa)[0])
