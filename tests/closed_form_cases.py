"""Closed-form known-answer cases shared by tests/test_oracle_cpu.py (pins oracle/lik_oracle.c) and
tests/test_closed_form_gpu.py (pins the CUDA engine directly): each case is
(name, exch, freqs, alpha, ncat, codes uint8 [n_tips, P], weights, desc, brlen, expected lnL)
with `expected` computed from tests/closed_form.py only."""
import itertools

import numpy as np

import closed_form as cf

F81_FREQS = np.array([0.1, 0.2, 0.3, 0.4])
HKY_FREQS = np.array([0.35, 0.15, 0.2, 0.3])


def _all_patterns(n_tips):
    return np.array(list(itertools.product(range(4), repeat=n_tips)), dtype=np.int64).T  # [n_tips, 4^n]


def _models():
    jc = ("JC69", np.ones(6), np.full(4, 0.25), lambda t: cf.p_jc69(t))
    f81 = ("F81", np.ones(6), F81_FREQS, lambda t: cf.p_f81(F81_FREQS, t))
    hky = ("HKY85", cf.hky_exch(4.3), HKY_FREQS, lambda t: cf.p_hky85(HKY_FREQS, 4.3, t))
    return [jc, f81, hky]


def cases():
    out = []
    rng = np.random.default_rng(20261020)
    for name, exch, freqs, pfun in _models():
        for alpha, ncat in ((1.0, 1), (0.7, 4), (2.5, 4)):
            rates = cf.gamma_mean_rates(alpha, ncat)
            # two taxa: every pattern, weights 1..16; the root sits anywhere on the single branch
            st = _all_patterns(2)
            w = np.arange(1, st.shape[1] + 1, dtype=float)
            for t1, t2 in ((0.05, 0.2), (0.25, 0.0), (1.7, 2.9)):
                want = cf.star_lnl(pfun, freqs, st, [t1 + t2, 0.0], w, rates)
                out.append(("%s two taxa a=%g R=%d t=%g+%g" % (name, alpha, ncat, t1, t2), exch, freqs, alpha, ncat,
                            (1 << st).astype(np.uint8), w, np.array([[0, 1]], dtype=np.int32),
                            np.array([[t1, t2]]), want))
            # three taxa: the unrooted tree IS a star; all 64 patterns
            st = _all_patterns(3)
            w = rng.integers(1, 9, st.shape[1]).astype(float)
            t0, t1, ta, tb = 0.11, 0.42, 0.07, 0.26
            want = cf.star_lnl(pfun, freqs, st, [t0, t1, ta + tb], w, rates)
            for a, b in ((ta, tb), (ta + tb, 0.0), (0.0, ta + tb)):  # pulley principle: same lnL wherever the root is
                out.append(("%s three taxa a=%g R=%d root %g+%g" % (name, alpha, ncat, a, b), exch, freqs, alpha, ncat,
                            (1 << st).astype(np.uint8), w, np.array([[0, 1], [3, 2]], dtype=np.int32),
                            np.array([[t0, t1], [a, b]]), want))
            # four taxa, internal branch of length zero = a four-pointed star; all 256 patterns
            st = _all_patterns(4)
            w = np.ones(st.shape[1])
            ts = [0.3, 0.05, 0.6, 0.12]
            want = cf.star_lnl(pfun, freqs, st, ts, w, rates)
            out.append(("%s four-star a=%g R=%d" % (name, alpha, ncat), exch, freqs, alpha, ncat,
                        (1 << st).astype(np.uint8), w, np.array([[0, 1], [2, 3], [4, 5]], dtype=np.int32),
                        np.array([[ts[0], ts[1]], [ts[2], ts[3]], [0.0, 0.0]]), want))
            # four taxa with an internal branch x: explicit double sum over the two inner states
            x = 0.23
            site = np.zeros(st.shape[1])
            for r in rates:
                P = [pfun(r * t) for t in ts]
                Px = pfun(r * x)
                for s in range(st.shape[1]):
                    left = P[0][:, st[0, s]] * P[1][:, st[1, s]]
                    right = P[2][:, st[2, s]] * P[3][:, st[3, s]]
                    site[s] += float(np.sum(freqs * left * (Px @ right))) / len(rates)
            out.append(("%s four taxa internal %g a=%g R=%d" % (name, x, alpha, ncat), exch, freqs, alpha, ncat,
                        (1 << st).astype(np.uint8), w, np.array([[0, 1], [2, 3], [4, 5]], dtype=np.int32),
                        np.array([[ts[0], ts[1]], [ts[2], ts[3]], [0.1, x - 0.1]]), float(np.sum(w * np.log(site)))))
        # saturation: infinitely long branches -> every site's likelihood is the product of its tips' frequencies
        st = _all_patterns(3)
        want = float(np.sum(np.log(freqs[st[0]] * freqs[st[1]] * freqs[st[2]])))
        out.append(("%s saturated" % name, exch, freqs, 1.0, 1, (1 << st).astype(np.uint8), np.ones(st.shape[1]),
                    np.array([[0, 1], [3, 2]], dtype=np.int32), np.array([[400.0, 400.0], [400.0, 400.0]]), want))
    return out
