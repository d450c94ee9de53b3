"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the CPU oracles.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; nothing under ``pylogeny_b200/``
does (tests/test_host_cpu.py::test_product_never_touches_the_oracle greps for it).

* ``fitch_cost``      -- plain-C restatement of /root/reference/pylogeny/fitch.cpp
                         (oracle/fitch_oracle.c); parity pinned by tests/golden/.
* ``ref_fitch_cost``  -- the reference's OWN fitch.cpp:6-272 compiled into
                         oracle/_ref/libfitch_ref.so (None if not built).
* ``lnl`` etc.        -- FP64 pruning restatement (oracle/lik_oracle.c); libpll is
                         absent, so likelihood parity is UNPINNED (see its header).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def build():
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_fitch_cost.restype = C.c_int
        L.oracle_fitch_cost.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64),
                                        C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        dp = C.POINTER(C.c_double)
        L.oracle_gamma_rates.argtypes = [C.c_double, C.c_int, dp]
        L.oracle_eigen.argtypes = [C.c_int, dp, dp, dp, dp, dp]
        L.oracle_pmatrix.argtypes = [C.c_int, dp, dp, dp, C.c_double, dp]
        for f in (L.oracle_lnl, L.oracle_lnl_bruteforce):
            f.restype = C.c_double
            f.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, dp, dp, dp, C.c_double, C.c_int,
                          C.c_int, C.c_void_p, dp]
        _LIB = L
    return _LIB


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libfitch_ref.so"))


def _ref():
    global _REF
    if _REF is None:
        L = C.CDLL(os.path.join(_HERE, "_ref", "libfitch_ref.so"))
        L.ref_fitch_cost.restype = C.c_int
        L.ref_fitch_cost.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_long), C.c_int,
                                     C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _REF = L
    return _REF


class FitchInputs:
    """Marshalled (profiles, weights, taxa) so many trees can be scored without re-marshalling."""

    def __init__(self, profiles, weights, taxa):
        self.n = len(profiles)
        self._p = [p.encode() for p in profiles]
        self.parr = (C.c_char_p * self.n)(*self._p)
        self.w64 = (C.c_int64 * self.n)(*[int(w) for w in weights])
        self.wl = (C.c_long * self.n)(*[int(w) for w in weights])
        names = list(taxa.keys())
        self.nt = len(names)
        self._n = [t.encode() for t in names]
        self.narr = (C.c_char_p * self.nt)(*self._n)
        self.i32 = (C.c_int32 * self.nt)(*[int(taxa[t]) for t in names])
        self.ii = (C.c_int * self.nt)(*[int(taxa[t]) for t in names])


def fitch_cost(newick, profiles=None, weights=None, taxa=None, inputs=None):
    """calculateCost(newick, profiles, weights, taxa) per fitch.cpp:274-334 (C restatement)."""
    x = inputs or FitchInputs(profiles, weights, taxa)
    out = C.c_int32(0)
    rc = _lib().oracle_fitch_cost(newick.encode(), x.n, x.parr, x.w64, x.nt, x.narr, x.i32, C.byref(out))
    if rc:
        raise ValueError("oracle_fitch_cost failed rc=%d" % rc)
    return out.value


def ref_fitch_cost(newick, profiles=None, weights=None, taxa=None, inputs=None):
    """The reference's own compiled core (oracle/_ref)."""
    x = inputs or FitchInputs(profiles, weights, taxa)
    out = C.c_int(0)
    rc = _ref().ref_fitch_cost(newick.encode(), x.n, x.parr, x.wl, x.nt, x.narr, x.ii, C.byref(out))
    if rc:
        raise ValueError("reference fitch core threw")
    return out.value


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def gamma_rates(alpha, ncat=4):
    out = np.zeros(ncat)
    _lib().oracle_gamma_rates(float(alpha), int(ncat), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def eigen(exch, freqs):
    K = len(freqs)
    e, ep = _d(exch)
    f, fp = _d(freqs)
    ev, U, Ui = np.zeros(K), np.zeros((K, K)), np.zeros((K, K))
    dp = C.POINTER(C.c_double)
    _lib().oracle_eigen(K, ep, fp, ev.ctypes.data_as(dp), U.ctypes.data_as(dp), Ui.ctypes.data_as(dp))
    return ev, U, Ui


def pmatrix(exch, freqs, t):
    ev, U, Ui = eigen(exch, freqs)
    K = len(freqs)
    P = np.zeros((K, K))
    dp = C.POINTER(C.c_double)
    _lib().oracle_pmatrix(K, ev.ctypes.data_as(dp), U.ctypes.data_as(dp), Ui.ctypes.data_as(dp), float(t),
                          P.ctypes.data_as(dp))
    return P


DNA_MASKS = np.arange(16, dtype=np.uint32)
AA_MASKS = np.array([1 << i for i in range(20)] + [(1 << 2) | (1 << 3), (1 << 5) | (1 << 6), (1 << 20) - 1],
                    dtype=np.uint32)  # 20 states (ARNDCQEGHILKMFPSTWYV), B = D|N, Z = Q|E, X/-/? = any


def lnl(codes, weights, exch, freqs, alpha, desc, brlen, ncat=4, masks=None, brute=False):
    """lnL of one tree. codes: uint8 [n_tips, P]; desc: int32 [n_inner, 2]; brlen: float64 [n_inner, 2]."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    n_tips, P = codes.shape
    K = len(freqs)
    if masks is None:
        masks = DNA_MASKS if K == 4 else AA_MASKS
    masks = np.ascontiguousarray(masks, dtype=np.uint32)
    w, wp = _d(weights)
    e, ep = _d(exch)
    f, fp = _d(freqs)
    desc = np.ascontiguousarray(desc, dtype=np.int32)
    b, bp = _d(brlen)
    fn = _lib().oracle_lnl_bruteforce if brute else _lib().oracle_lnl
    return fn(K, n_tips, P, codes.ctypes.data, masks.ctypes.data, wp, ep, fp, float(alpha), int(ncat),
              desc.shape[0], desc.ctypes.data, bp)
