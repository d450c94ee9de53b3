"""ctypes binding of include/pylogeny_b200.h.  Fails loudly: there is no CPU fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PLG_LIB_PATH") or os.path.join(_HERE, "libpylogeny_b200.so")  # override: tuning sweeps

OK, ERR_ARG, ERR_PARSE, ERR_LABEL, ERR_CUDA, ERR_STATES, ERR_UNSUPPORTED, ERR_IO = range(8)
MOVE_SPR, MOVE_NNI, MOVE_TBR, MOVES_ALL = 1, 2, 3, 4


def MOVE_SPR_WITHIN(radius):
    """SPR moves regrafting at most `radius` branches away (PLG_MOVE_SPR_WITHIN; 0 = unlimited)."""
    if radius < 0:
        raise ValueError("radius must be >= 0")
    return MOVE_SPR | (int(radius) << 8)


DATA_DNA, DATA_AA = 4, 20
OUT_HOST, OUT_DEVICE = 0, 1

_lib = None


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "pylogeny_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (python -m pylogeny_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -m pylogeny_b200.build` (needs nvcc); "
                              "pylogeny_b200 has no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        _declare(L)
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise EngineError(rc, lib().plg_last_error().decode(errors="replace"))


def _set(L, name, attr, value):
    """Declare one entry point; a library built before the entry existed (tuning variants) still loads,
    and calling the missing entry raises AttributeError."""
    fn = getattr(L, name, None)
    if fn is not None:
        setattr(fn, attr, value)


def _declare(L):
    i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    strs = C.POINTER(C.c_char_p)
    vp = C.c_void_p
    _set(L, 'plg_last_error', 'restype', C.c_char_p)
    _set(L, 'plg_last_error', 'argtypes', [])
    _set(L, 'plg_version', 'restype', C.c_int)
    _set(L, 'plg_launch_count', 'restype', C.c_int64)
    _set(L, 'plg_device_count', 'argtypes', [C.POINTER(C.c_int)])
    _set(L, 'plg_set_device', 'argtypes', [C.c_int])
    _set(L, 'plg_fitch_aln_create', 'argtypes', [C.c_int64, strs, i64p, C.POINTER(vp)])
    _set(L, 'plg_fitch_aln_create_dense', 'argtypes', [C.c_int64, C.c_int, vp, vp, C.POINTER(vp)])
    _set(L, 'plg_fitch_aln_destroy', 'argtypes', [vp])
    _set(L, 'plg_fitch_aln_destroy', 'restype', None)
    _set(L, 'plg_fitch_aln_info', 'argtypes', [vp, i64p, C.POINTER(C.c_int), C.POINTER(C.c_int)])
    _set(L, 'plg_fitch_cost', 'argtypes', [C.c_char_p, C.c_int64, strs, i64p, C.c_int, strs, i32p, i32p])
    _set(L, 'plg_fitch_score_newicks', 'argtypes', [vp, C.c_int64, strs, C.c_int, strs, i32p, vp, vp])
    _set(L, 'plg_fitch_score_trees', 'argtypes', [vp, C.c_int64, C.c_int, vp, vp, vp, vp])
    dp = C.POINTER(C.c_double)
    _set(L, 'plg_lik_aln_create', 'argtypes', [C.c_int, C.c_int, C.c_int64, vp, vp, C.POINTER(vp)])
    _set(L, 'plg_lik_aln_destroy', 'argtypes', [vp])
    _set(L, 'plg_lik_aln_destroy', 'restype', None)
    _set(L, 'plg_model_create', 'argtypes', [C.c_int, dp, dp, C.c_double, C.c_int, C.POINTER(vp)])
    _set(L, 'plg_protein_model', 'argtypes', [C.c_char_p, dp, dp])
    _set(L, 'plg_fp64_peak', 'argtypes', [C.c_int, dp])
    _set(L, 'plg_model_destroy', 'argtypes', [vp])
    _set(L, 'plg_model_destroy', 'restype', None)
    _set(L, 'plg_model_gamma_rates', 'argtypes', [vp, dp])
    _set(L, 'plg_model_pmatrix', 'argtypes', [vp, C.c_double, dp])
    _set(L, 'plg_lik_score_trees', 'argtypes', [vp, vp, C.c_int64, C.c_int, vp, vp, vp, vp, vp])
    u64p = C.POINTER(C.c_uint64)
    _set(L, 'plg_neighbourhood_size', 'argtypes', [C.c_int, C.c_int, vp, i64p])
    _set(L, 'plg_neighbourhood_trees', 'argtypes', [C.c_int, C.c_int, vp, vp, vp, C.c_int64, vp, vp, vp, vp])
    _set(L, 'plg_topology_hash', 'argtypes', [C.c_int, vp, vp, u64p])
    _set(L, 'plg_hostpool_selftest', 'argtypes', [C.c_int, C.c_int, C.POINTER(C.c_int)])
    _set(L, 'plg_neighbourhood_all_moves', 'argtypes', [C.c_int, C.c_int, vp, C.POINTER(C.c_int64), vp])
    _set(L, 'plg_fitch_score_neighbourhood', 'argtypes', [vp, C.c_int, C.c_int, vp, vp, C.c_int64, C.c_int64, i64p, vp, vp, vp])
    _set(L, 'plg_lik_score_neighbourhood', 'argtypes',
         [vp, vp, C.c_int, C.c_int, vp, vp, vp, C.c_int64, C.c_int64, i64p, vp, vp, vp])
    _set(L, 'plg_fitch_score_trees_shard', 'argtypes', [vp, C.c_int64, C.c_int, vp, vp, C.c_int64, C.c_int64, vp, C.c_int, vp])
    _set(L, 'plg_lik_score_trees_shard', 'argtypes',
         [vp, vp, C.c_int64, C.c_int, vp, vp, vp, C.c_int64, C.c_int64, vp, C.c_int, vp])
    _set(L, 'plg_fitch_score_neighbourhood_out', 'argtypes',
         [vp, C.c_int, C.c_int, vp, vp, C.c_int64, C.c_int64, i64p, vp, vp, C.c_int, vp])
    _set(L, 'plg_lik_score_neighbourhood_out', 'argtypes',
         [vp, vp, C.c_int, C.c_int, vp, vp, vp, C.c_int64, C.c_int64, i64p, vp, vp, C.c_int, vp])
    _set(L, 'plg_profile_enable', 'argtypes', [C.c_int])
    _set(L, 'plg_profile_read', 'argtypes', [C.c_int, dp, dp, dp, i64p, dp])
    _set(L, 'plg_pll_new', 'argtypes', [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(vp)])
    _set(L, 'plg_pll_destroy', 'argtypes', [vp])
    _set(L, 'plg_pll_loglik', 'argtypes', [vp, dp])
    _set(L, 'plg_pll_evaluate', 'argtypes', [vp, dp])
    _set(L, 'plg_pll_newick', 'argtypes', [vp, C.c_char_p, C.POINTER(C.c_size_t)])
    _set(L, 'plg_pll_moves_compute', 'argtypes', [vp, C.c_int, C.c_int, i64p, C.POINTER(C.c_size_t)])
    _set(L, 'plg_pll_moves_fetch', 'argtypes', [vp, vp, vp, vp, C.c_size_t])
    _set(L, 'plg_pll_info', 'argtypes', [vp, C.POINTER(C.c_int), i64p, dp, dp])
    _set(L, 'plg_pll_loglik_batch', 'argtypes', [C.c_char_p, C.c_char_p, C.c_int64, strs, C.c_int, vp, C.POINTER(C.c_size_t)])
    _set(L, 'plg_pll_batch_newicks', 'argtypes', [C.c_char_p, C.c_size_t])
