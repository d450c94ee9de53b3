"""The CUDA engine against closed-form likelihoods (tests/closed_form.py: JC69 / F81 / HKY85
transition probabilities as formulas, Yang-1994 Gamma rates from scipy, star trees as explicit
sums) -- numbers independent of this repository's eigendecomposition, pruning order and scaling.
Reference path pinned: libpllWrapper.c:77-78 (pllEvaluateLikelihood at given branch lengths)."""
import numpy as np
import pytest

import closed_form_cases as cc

pytestmark = pytest.mark.gpu


def test_engine_equals_closed_forms():
    from pylogeny_b200 import likelihood
    worst = 0.0
    for name, exch, freqs, alpha, ncat, codes, w, desc, br, want in cc.cases():
        la = likelihood.LikAlignment(codes, w)
        m = likelihood.Model(exch, freqs, alpha, ncat)
        got = float(la.score_trees(m, desc[None], br[None])[0])
        la.close(); m.close()
        rel = abs(got - want) / abs(want)
        worst = max(worst, rel)
        assert rel <= 1e-11, (name, got, want)
    assert worst <= 1e-11


def test_engine_neighbourhood_of_a_quartet_closed_form():
    """The SPR/NNI neighbourhood of a four-taxon tree (2 neighbours) by edge insertion: each
    neighbour is another quartet whose likelihood is the explicit sum over all inner-state assignments, with the branch lengths the reference's move gives it (rearrangement.py:466-497)."""
    import closed_form as cf
    from pylogeny_b200 import likelihood, neighbourhood as nbh
    import itertools
    st = np.array(list(itertools.product(range(4), repeat=4)), dtype=np.int64).T
    freqs, kappa, alpha = cc.HKY_FREQS, 4.3, 0.7
    rates = cf.gamma_mean_rates(alpha, 4)
    desc = np.array([[0, 1], [2, 3], [4, 5]], dtype=np.int32)
    br = np.array([[0.3, 0.05], [0.6, 0.12], [0.1, 0.13]])
    la = likelihood.LikAlignment((1 << st).astype(np.uint8), np.ones(st.shape[1]))
    m = likelihood.Model(cf.hky_exch(kappa), freqs, alpha, 4)
    _, lnl = nbh.score_likelihood(la, m, desc, br)
    descs, brs, _ = nbh.neighbour_trees(desc, br)
    assert len(lnl) == 2
    pfun = lambda t: cf.p_hky85(freqs, kappa, t)
    for k in range(len(lnl)):
        want = cf.tree_lnl_sum_over_states(pfun, freqs, descs[k], brs[k], st, None, rates)
        assert abs(lnl[k] - want) <= 1e-11 * abs(want), (k, lnl[k], want)
    la.close(); m.close()
