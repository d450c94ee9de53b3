"""Drop-in for the reference's `libpllWrapper` extension module (libpllWrapper.c:323-339).

    new(alignment_file, newick, partition_file) -> handle      (:97-171)
    destroy(handle) -> None                                    (:173-193)
    getLogLikelihood(handle) -> float   (mutates branch lengths, :195-215)
    getNewickString(handle) -> str                             (:217-239)

over the C-ABI `plg_pll_*` (include/pylogeny_b200.h).  Error behaviour follows the reference:
IOError for unparsable inputs (:141-146, :153-156, :211), ReferenceError for a bad handle (:178).
    getSPRMovesInDistance / getNNIMovesInDistance / getAllMovesInDistance(handle, maxdist)
        -> [("SPR"|"NNI", lnL, newick)]                        (:241-319, list built at :36-93)

The three move enumerators have no Python caller in the reference.  Here they
return EVERY distinct neighbour of the handle's current tree within `maxdist` branches (libpll
searches only from nodep[1..n+1], :254, and caps the list at 4(n-3)(n-2)+(n-1), :86-88), scored at
the handle's current branch lengths with the neighbour's lengths set by the reference's own SPR
rule (rearrangement.py:466-497), lnL rounded through float like :54, best first.  The bulk form
without strings is pylogeny_b200.neighbourhood.score_likelihood.
"""
import numpy as np

import ctypes as C

from . import _lib

__all__ = ["new", "destroy", "getLogLikelihood", "getNewickString", "evaluate",
           "getSPRMovesInDistance", "getNNIMovesInDistance", "getAllMovesInDistance"]


class _Handle(object):
    __slots__ = ("ptr",)

    def __init__(self, ptr):
        self.ptr = ptr


def _ptr(handle):
    if not isinstance(handle, _Handle) or not handle.ptr:
        raise ReferenceError("Could not parse pointer.")
    return handle.ptr


def _raise(rc):
    msg = _lib.lib().plg_last_error().decode(errors="replace")
    if rc in (_lib.ERR_IO, _lib.ERR_PARSE, _lib.ERR_LABEL):
        raise IOError(msg)
    raise _lib.EngineError(rc, msg)


def new(alignf, newick, partf):
    """Initialize a problem with an alignment, Newick tree, and partition model."""
    if not all(isinstance(x, str) for x in (alignf, newick, partf)):
        raise IOError("Need alignment file, newick string, partition file as strings.")
    h = C.c_void_p()
    rc = _lib.lib().plg_pll_new(alignf.encode(), newick.encode(), partf.encode(), C.byref(h))
    if rc:
        _raise(rc)
    return _Handle(h)


def destroy(handle):
    """Destroy an input problem and deallocate all resources."""
    p = _ptr(handle)
    _lib.lib().plg_pll_destroy(p)
    handle.ptr = None


def getLogLikelihood(handle):
    """Score a tree and acquire its log-likelihood for a given input problem."""
    out = C.c_double()
    rc = _lib.lib().plg_pll_loglik(_ptr(handle), C.byref(out))
    if rc:
        _raise(rc)
    if not out.value:
        raise IOError("Unable to evaluate tree for likelihood.")
    return out.value


def getLogLikelihoodBatch(alignf, newicks, partf, passes=3):
    """Extension: what new() + getLogLikelihood() + getNewickString() + destroy() give for every
    tree of `newicks` (default branch lengths, `passes` smoothing passes, likelihood), all trees
    advancing in lockstep on the device.  Returns (list of lnL, list of optimised Newick strings)."""
    import numpy as np
    L = _lib.lib()
    n = len(newicks)
    arr = (C.c_char_p * max(n, 1))(*[s.encode() for s in newicks])
    lnl = np.zeros(n, dtype=np.float64)
    nbytes = C.c_size_t(0)
    rc = L.plg_pll_loglik_batch(alignf.encode(), partf.encode(), n, arr, int(passes), lnl.ctypes.data, C.byref(nbytes))
    if rc:
        _raise(rc)
    buf = C.create_string_buffer(max(1, nbytes.value))
    rc = L.plg_pll_batch_newicks(buf, nbytes.value)
    if rc:
        _raise(rc)
    out = buf.raw[:nbytes.value].split(b"\0")[:n]
    return lnl.tolist(), [b.decode() for b in out]


def evaluate(handle):
    """lnL at the current branch lengths, no optimisation (extension; pllEvaluateLikelihood)."""
    out = C.c_double()
    rc = _lib.lib().plg_pll_evaluate(_ptr(handle), C.byref(out))
    if rc:
        _raise(rc)
    return out.value


def getNewickString(handle):
    """Get the Newick string for the current initialized problem."""
    n = C.c_size_t(0)
    _lib.lib().plg_pll_newick(_ptr(handle), None, C.byref(n))
    buf = C.create_string_buffer(n.value)
    rc = _lib.lib().plg_pll_newick(_ptr(handle), buf, C.byref(n))
    if rc:
        raise IOError("Unable to acquire Newick string for tree.")
    return buf.value.decode()


def _moves(handle, which, maxdist):
    if not isinstance(maxdist, int) or isinstance(maxdist, bool):
        raise ReferenceError("Could not parse pointer, integer.")  # libpllWrapper.c:246-248
    p = _ptr(handle)
    L = _lib.lib()
    n, nbytes = C.c_int64(), C.c_size_t()
    rc = L.plg_pll_moves_compute(p, which, maxdist, C.byref(n), C.byref(nbytes))
    if rc:
        _raise(rc)
    if n.value == 0:
        return []
    kinds = np.zeros(n.value, dtype=np.int32)
    lnl = np.zeros(n.value, dtype=np.float32)
    buf = C.create_string_buffer(nbytes.value)
    rc = L.plg_pll_moves_fetch(p, kinds.ctypes.data, lnl.ctypes.data, buf, nbytes.value)
    if rc:
        _raise(rc)
    strings = buf.raw[:-1].split(b"\0")
    return [("NNI" if k == _lib.MOVE_NNI else "SPR", float(v), s.decode())
            for k, v, s in zip(kinds.tolist(), lnl.tolist(), strings)]


def getSPRMovesInDistance(handle, maxdist):
    """Get all SPR moves in a given radius of the current tree."""
    return _moves(handle, _lib.MOVE_SPR, maxdist)


def getNNIMovesInDistance(handle, maxdist):
    """Get all NNI moves in a given radius of the current tree."""
    return _moves(handle, _lib.MOVE_NNI, maxdist)


def getAllMovesInDistance(handle, maxdist):
    """Get all NNI and SPR moves in a given radius of the current tree."""
    return _moves(handle, _lib.MOVES_ALL, maxdist)
