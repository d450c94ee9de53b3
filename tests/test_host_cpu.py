"""CPU-side checks: host logic (profile_set, alignment reader), the C-ABI boundary (library loads,
every declared symbol exported) and product/oracle separation.  No GPU compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def _alignment():
    from pylogeny_b200 import alignment
    al = load_golden("al_alignment.json")
    return alignment.alignment(names=al["taxa"], seqs=al["seqs"])


def test_profile_set_matches_reference_golden(al_profiles):
    """Same profiles, first-occurrence order, weights and taxa dict as reference parsimony.profile_set."""
    from pylogeny_b200 import parsimony
    ps = parsimony.profile_set(_alignment())
    assert [str(p) for p in ps.profiles] == al_profiles["profiles"]
    assert ps.weights == al_profiles["weights"]
    assert ps.taxa == al_profiles["taxa"]
    assert ps.numSites == al_profiles["n_sites"] == len(ps.sites) == sum(ps.weights)
    assert len(ps) == 142
    assert [(str(p), w) for p, w in zip(ps.profiles[:5], ps.weights[:5])] == [
        ('00010100', 38), ('00010120', 2), ('00012122', 3), ('00012102', 1), ('00010000', 35)]


def test_profile_set_generic_path_equals_vectorised(al_profiles):
    """An alignment object without as_matrix() (only the reference's p4-style surface) gives the same set."""
    from pylogeny_b200 import parsimony
    al = _alignment()

    class Plain(object):
        data = al.data
        getSize = staticmethod(al.getSize)
        getTaxa = staticmethod(al.getTaxa)

    ps = parsimony.profile_set(Plain())
    assert [str(p) for p in ps.profiles] == al_profiles["profiles"] and ps.weights == al_profiles["weights"]
    assert ps.getForTaxa(1, al.getTaxa()[7]) == '0'
    assert ps.get(2) == ps.profiles[2] and ps.weight(2) == 3


def test_profile_set_many_symbols():
    """>= 10 distinct symbols in a column: str(idx) turns multi-char exactly as parsimony.py:144 does."""
    from pylogeny_b200 import alignment, parsimony
    seqs = [c for c in "ABCDEFGHIJKLM"]
    al = alignment.alignment(names=["t%d" % i for i in range(13)], seqs=seqs)
    ps = parsimony.profile_set(al)
    assert str(ps.profiles[0]) == "0123456789101112"


def test_phylip_friendly_alignment(tmp_path):
    from pylogeny_b200 import alignment
    g = load_golden("al_alignment.json")
    fa = tmp_path / "a.fasta"
    fa.write_text("".join(">%s\n%s\n" % (n, s) for n, s in zip(g["taxa"], g["seqs"])))
    al = alignment.phylipFriendlyAlignment(str(fa))
    assert al.getTaxa() == ["T%d" % i for i in range(8)]
    assert al.getProperName("T3") == g["taxa"][3]
    assert al.getSize() == 1431 and al.getNumSeqs() == 8 and al.getDataType() == 'dna'
    ph = al.getPhylip()
    back = alignment.alignment(ph)
    assert back.getTaxa() == al.getTaxa() and back.toStrList() == al.toStrList()
    al.close()
    assert not os.path.exists(ph)


def test_library_exports_every_declared_symbol():
    from pylogeny_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pylogeny_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(plg_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 10
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert _lib.lib().plg_version() >= 100


def test_no_gpu_fails_loudly():
    """Without a CUDA device the compute entry points report PLG_ERR_CUDA instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pylogeny_b200 import _lib, fitch
    with pytest.raises(_lib.EngineError) as e:
        fitch.calculateCost("(A,B);", ["01"], [1], {"A": 0, "B": 1})
    assert e.value.code == _lib.ERR_CUDA


def test_product_never_touches_the_oracle():
    """Only tests/, bench.py's CPU legs and __graft_entry__.smoke() may use oracle/."""
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "pylogeny_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dp, fn), errors="replace").read()
                # imports, dlopen targets or paths into oracle/ (comments that merely say "oracle" are fine)
                if re.search(r"(import\s+oracle|from\s+oracle|liboracle|libfitch_ref|oracle/|oracle\.)", src):
                    bad.append(fn)
    assert not bad, bad


def test_newick_errors_are_reported_not_thrown():
    from pylogeny_b200 import _lib
    import ctypes as C
    out = C.c_int32()
    rc = _lib.lib().plg_fitch_cost(b"((A,B);", 0, None, None, 0, None, None, C.byref(out))
    assert rc == _lib.ERR_PARSE and b"Newick" in _lib.lib().plg_last_error()
    names = (C.c_char_p * 1)(b"A")
    idx = (C.c_int32 * 1)(0)
    rc = _lib.lib().plg_fitch_cost(b"(A,B);", 0, None, None, 1, names, idx, C.byref(out))
    assert rc == _lib.ERR_LABEL


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one line on
    stdout, valid JSON with the contract's keys; everything else goes to stderr."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PLG_BENCH_REF_SECONDS="1.5")
    for workload, kinds, starts in ((None, ("port",), "north_star target shape"), ("fitch_target", ("reference", "port"), "BASELINE configs[2]"),
                                    ("aa_c5", ("port",), "BASELINE configs[4]")):
        cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"]
        if workload:
            cmd += ["--workload", workload]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=300, text=True)
        assert r.returncode == 0, r.stderr[-500:]
        lines = [l for l in r.stdout.splitlines() if l.strip()]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["unit"] == "tree-sites/s" and d["value"] > 0 and d["higher_is_better"] is True
        assert d["cpu_baseline"]["kind"] in kinds and d["cpu_baseline"]["cores"] >= 1
        assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert d["config"]["workload"].startswith(starts)


def test_bench_reference_arm_is_product_free():
    """The CPU arm scores trees the REFERENCE's rearrangement code enumerated (committed fixture) and never
    loads the product: no pylogeny_b200 module, no libpylogeny_b200.so in the process."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, json; sys.argv = ['bench.py']; sys.path.insert(0, %r); import bench\n"
            "cb, dt = bench.cpu_sample('lnl_c2', 0.5)\n"
            "maps = open('/proc/self/maps').read()\n"
            "print(json.dumps({'mods': [m for m in sys.modules if m.startswith('pylogeny_b200')], "
            "'so': 'libpylogeny_b200' in maps, 'kind': cb['kind'], 'v': cb['value']}))" % root)
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    import json
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["mods"] == [] and d["so"] is False and d["kind"] == "port" and d["v"] > 0


def test_protein_tables():
    """The shipped WAG and JTT tables (csrc/protein_models.cpp, typed from memory of PAML's wag.dat /
    jones.dat): what can be checked without the published files -- shape, positivity, frequencies
    summing to 1 within the files' rounding, landmark entries, the dominant exchanges of either
    paper, and that the rate matrix built from them has the table's frequencies as its stationary
    distribution (detailed balance).  See the file header for what stays unchecked."""
    from pylogeny_b200 import likelihood
    order = likelihood.AA_ORDER
    idx = {c: i for i, c in enumerate(order)}
    top_expected = {"WAG": [("I", "V"), ("F", "Y"), ("D", "E"), ("D", "N"), ("E", "Q"), ("K", "R")],
                    "JTT": [("I", "V"), ("D", "E"), ("K", "R"), ("H", "Q"), ("H", "Y"), ("F", "Y")]}
    landmarks = {"WAG": {("A", "R"): 0.551571, ("D", "N"): 5.429420, ("I", "V"): 7.821300, ("C", "D"): 0.0302949},
                 "JTT": {("A", "R"): 58.0, ("D", "E"): 767.0, ("I", "V"): 961.0, ("C", "E"): 5.0}}
    for name in ("WAG", "JTT", "wag"):
        exch, freqs = likelihood.protein_model(name)
        assert exch.shape == (190,) and freqs.shape == (20,) and np.all(exch > 0) and np.all(freqs > 0)
        assert abs(freqs.sum() - 1.0) < 2e-6
        R = np.zeros((20, 20))
        R[np.triu_indices(20, 1)] = exch
        R = R + R.T
        key = name.upper()
        for (a, b), v in landmarks[key].items():
            assert R[idx[a], idx[b]] == v
        ranked = sorted(((R[i, j], order[i], order[j]) for i in range(20) for j in range(i + 1, 20)), reverse=True)
        top = {tuple(sorted((a, b))) for _, a, b in ranked[:12]}
        for pair in top_expected[key]:
            assert tuple(sorted(pair)) in top, (name, pair, ranked[:12])
        Q = R * freqs[None, :]
        np.fill_diagonal(Q, -Q.sum(1))
        np.testing.assert_allclose(freqs @ Q, 0.0, atol=1e-12)                        # stationary
        np.testing.assert_allclose(freqs[:, None] * Q, (freqs[:, None] * Q).T, atol=1e-15)  # reversible
    with pytest.raises(Exception):
        likelihood.protein_model("LG")


def test_host_pool_survives_a_failing_slice():
    """A planner slice that fails on a pool worker (or on the caller) comes back as the call's status
    code after all other slices ran; the pool then serves the next call (csrc/hostpool.h)."""
    import ctypes as C
    from pylogeny_b200 import _lib
    L = _lib.lib()
    ran = C.c_int()
    for n, fail_at in ((8, 5), (8, 0), (64, 63), (3, 1)):
        assert L.plg_hostpool_selftest(n, fail_at, C.byref(ran)) == _lib.ERR_ARG
        assert ran.value == n - 1 and b"failed on request" in L.plg_last_error()
        assert L.plg_hostpool_selftest(n, -1, C.byref(ran)) == 0 and ran.value == n
