"""Batched Felsenstein-pruning log-likelihood on the B200 engine (the path behind pll.py /
libpllWrapper.py): device-resident tip codes, a reversible model with Gamma-4 rates, and
descriptor batches of trees scored in one C-ABI call (include/pylogeny_b200.h: plg_lik_*)."""
import ctypes as C

import numpy as np

from . import _lib

DNA_CODE = {c: m for c, m in zip("ACGTUMRWSYKVHDBN-?XO.",
                                 [1, 2, 4, 8, 8, 3, 5, 9, 6, 10, 12, 7, 11, 13, 14, 15, 15, 15, 15, 15, 15])}
AA_ORDER = "ARNDCQEGHILKMFPSTWYV"
AA_CODE = dict({c: i for i, c in enumerate(AA_ORDER)}, B=20, Z=21, X=22, **{"-": 22, "?": 22, "*": 22, ".": 22})


def encode_sequences(seqs, datatype):
    """List of equal-length strings -> uint8 [n_rows, n_sites] tip codes (DNA: 4-bit masks; AA: 0..22)."""
    table = np.full(256, 15 if datatype == _lib.DATA_DNA else 22, dtype=np.uint8)
    for ch, code in (DNA_CODE if datatype == _lib.DATA_DNA else AA_CODE).items():
        table[ord(ch)] = code
        table[ord(ch.lower())] = code
    raw = np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8).reshape(len(seqs), -1)
    return table[raw]


def compress_columns(codes):
    """Identical columns -> (unique codes [n_rows, P], weights [P]) in first-occurrence order."""
    cols = np.ascontiguousarray(codes.T)
    void = cols.view(np.dtype((np.void, cols.shape[1]))).ravel()
    _, first, counts = np.unique(void, return_index=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    return np.ascontiguousarray(cols[first[order]].T), counts[order].astype(np.float64)


def protein_model(name):
    """(exch [190], freqs [20]) of a built-in empirical protein model: "WAG" or "JTT" (the names a
    partition file carries, pll.py:102-115).  Host lookup; table provenance: csrc/protein_models.cpp."""
    exch, freqs = np.zeros(190), np.zeros(20)
    dp = C.POINTER(C.c_double)
    _lib.check(_lib.lib().plg_protein_model(name.encode(), exch.ctypes.data_as(dp), freqs.ctypes.data_as(dp)))
    return exch, freqs


class Model(object):
    """Reversible substitution model + discrete Gamma (plg_model)."""

    def __init__(self, exch, freqs, alpha=1.0, n_cat=4):
        e = np.ascontiguousarray(exch, dtype=np.float64)
        f = np.ascontiguousarray(freqs, dtype=np.float64)
        K = f.shape[0]
        if e.shape[0] != K * (K - 1) // 2:
            raise ValueError("need K(K-1)/2 exchangeabilities")
        self.K, self.n_cat = K, n_cat
        self._h = C.c_void_p()
        dp = C.POINTER(C.c_double)
        _lib.check(_lib.lib().plg_model_create(K, e.ctypes.data_as(dp), f.ctypes.data_as(dp), float(alpha),
                                               int(n_cat), C.byref(self._h)))

    def gamma_rates(self):
        out = np.zeros(self.n_cat)
        _lib.check(_lib.lib().plg_model_gamma_rates(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def pmatrix(self, t):
        out = np.zeros((self.K, self.K))
        _lib.check(_lib.lib().plg_model_pmatrix(self._h, float(t), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().plg_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LikAlignment(object):
    """Device-resident tip codes + pattern weights (plg_lik_aln)."""

    def __init__(self, codes, weights, datatype=_lib.DATA_DNA):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if codes.ndim != 2 or codes.shape[1] != w.shape[0]:
            raise ValueError("codes must be [n_rows, n_patterns] with one weight per pattern")
        self.n_rows, self.n_patterns = codes.shape
        self.datatype = datatype
        self._h = C.c_void_p()
        _lib.check(_lib.lib().plg_lik_aln_create(int(datatype), self.n_rows, self.n_patterns, codes.ctypes.data,
                                                 w.ctypes.data, C.byref(self._h)))

    def score_trees(self, model, desc, brlen, tip_rows=None):
        """lnL of each tree: desc int32 [n_trees, n_tips-1, 2], brlen float64 same shape."""
        desc = np.ascontiguousarray(desc, dtype=np.int32)
        brlen = np.ascontiguousarray(brlen, dtype=np.float64)
        if desc.ndim == 2:
            desc, brlen = desc[None], brlen[None]
        if desc.shape != brlen.shape or desc.shape[2] != 2:
            raise ValueError("desc and brlen must both be [n_trees, n_tips-1, 2]")
        n_trees, ni, _ = desc.shape
        rows = np.ascontiguousarray(tip_rows, dtype=np.int32) if tip_rows is not None else None
        out = np.zeros(max(n_trees, 1))
        _lib.check(_lib.lib().plg_lik_score_trees(self._h, model._h, n_trees, ni + 1, desc.ctypes.data,
                                                  brlen.ctypes.data, rows.ctypes.data if rows is not None else None,
                                                  out.ctypes.data, None))
        return out[:n_trees]

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().plg_lik_aln_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
