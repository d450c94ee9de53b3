"""Test helpers: seeded random trees / alignments shared by CPU and GPU tests."""
import numpy as np


def random_desc(rng, n_tips):
    """Random-join rooted binary tree as a post-order descriptor int32 [n_tips-1, 2]."""
    nodes = list(range(n_tips))
    desc = []
    nxt = n_tips
    while len(nodes) > 1:
        a = nodes.pop(int(rng.integers(len(nodes))))
        b = nodes.pop(int(rng.integers(len(nodes))))
        desc.append((a, b))
        nodes.append(nxt)
        nxt += 1
    return np.array(desc, dtype=np.int32)


def desc_to_newick(desc, labels, brlen=None):
    n = len(labels)
    s = {}
    for i, (a, b) in enumerate(desc):
        def sub(c, k):
            t = labels[c] if c < n else s[c]
            if brlen is not None:
                t += ":%r" % float(brlen[i][k])
            return t
        s[n + i] = "(%s,%s)" % (sub(int(a), 0), sub(int(b), 1))
    return s[n + len(desc) - 1] + ";"


def random_kary_newick(rng, labels, multif=0.4):
    nodes = list(labels)
    rng.shuffle(nodes)
    while len(nodes) > 1:
        k = 2
        while k < len(nodes) and rng.random() < multif:
            k += 1
        picks = [nodes.pop(int(rng.integers(len(nodes)))) for _ in range(k)]
        nodes.append("(" + ",".join(picks) + ")")
    return nodes[0] + ";"


def random_states(rng, n_rows, P, n_states=4, gap_frac=0.0):
    """uint8 [P, n_rows] of digit characters '0'.. (fitch profile alphabet), canonicalised per pattern."""
    raw = rng.integers(0, n_states, (P, n_rows))
    if gap_frac:
        raw = np.where(rng.random((P, n_rows)) < gap_frac, n_states, raw)
    out = np.empty_like(raw)
    for p in range(P):  # first-appearance relabel so profiles look like parsimony.py:132-145 output
        m = {}
        for r in range(n_rows):
            out[p, r] = m.setdefault(int(raw[p, r]), len(m))
    return (out + 48).astype(np.uint8)


def states_to_profiles(states):
    return [bytes(row).decode("ascii") for row in states]


def simulate_dna(rng, desc, brlen, n_sites, exch, freqs):
    """Evolve ACGT sequences down a rooted descriptor tree with one rate category (test data with signal)."""
    import oracle
    n = desc.shape[0] + 1
    root = n + len(desc) - 1
    states = {root: rng.choice(4, n_sites, p=freqs)}
    for i in range(len(desc) - 1, -1, -1):
        for k, c in enumerate(desc[i]):
            P = oracle.pmatrix(exch, freqs, float(brlen[i][k]))
            P = np.clip(P, 0, None)
            P /= P.sum(1, keepdims=True)
            cum = P.cumsum(1)
            u = rng.random(n_sites)
            states[int(c)] = (u[:, None] > cum[states[n + i]]).sum(1).clip(0, 3)
    return ["".join("ACGT"[s] for s in states[t]) for t in range(n)]
