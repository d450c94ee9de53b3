#!/usr/bin/env python
"""Generate tests/golden/bench_ref_trees.npz: the sample trees bench.py's reference arm scores on the
CPU, enumerated and materialised by the REFERENCE's own rearrangement code -- so that arm touches
nothing of pylogeny_b200 (both arms scored trees from the product's enumerator).

Run in the build container only (needs /root/reference):  python tests/golden/make_bench_fixture.py

For each neighbourhood workload of bench.py (same seeds, same base trees) the reference's
`topology.iterSPRForBranch` (rearrangement.py:542-567) is walked over branches spread across the
tree, `rearrangement.toTree()` (:116-126) materialises each sampled move with the reference's own
branch-length rule (:466-497), and the resulting Newick string is stored as a post-order descriptor
+ branch lengths (the form the CPU oracle takes).  For the protein TBR workload -- the reference
only names TBR (rearrangement.py:27) and has no enumerator -- neighbours are made here by textbook
bisection + reconnection in plain Python.  Arrays only; /root/reference does not travel."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402  (seeds and base trees of the workloads; imports no product code at module level)
import refshim  # noqa: E402


def newick_to_arrays(s, n):
    """Newick with T<i> labels -> (desc int16 [n-1, 2], brlen float64 [n-1, 2]); tip id = i of T<i>.
    A trifurcating root is resolved with a zero-length branch."""
    s = s.strip().rstrip(";")
    pos = [0]
    desc, brl = [], []

    def length():
        if pos[0] < len(s) and s[pos[0]] == ":":
            j = pos[0] + 1
            while j < len(s) and s[j] not in ",)":
                j += 1
            v = float(s[pos[0] + 1:j])
            pos[0] = j
            return v
        return 0.0

    def sub():
        if s[pos[0]] == "(":
            pos[0] += 1
            kids = [sub()]
            while s[pos[0]] == ",":
                pos[0] += 1
                kids.append(sub())
            assert s[pos[0]] == ")"
            pos[0] += 1
            while pos[0] < len(s) and s[pos[0]] not in ",:)":
                pos[0] += 1
            ln = length()
            node, nl = kids[0]
            for k, kl in kids[1:]:
                desc.append((node, k))
                brl.append((nl, kl))
                node, nl = n + len(desc) - 1, 0.0
            return node, ln
        j = pos[0]
        while s[j] not in ",:)":
            j += 1
        lab = s[pos[0]:j]
        pos[0] = j
        return int(lab[1:]), length()

    sub()
    assert len(desc) == n - 1
    return np.array(desc, dtype=np.int16), np.array(brl, dtype=np.float64)


def arrays_to_newick(desc, brlen, n):
    text = {}
    for i, (a, b) in enumerate(desc):
        parts = []
        for k, c in enumerate((int(a), int(b))):
            parts.append((("T%d" % c) if c < n else text[c]) + ":%r" % float(brlen[i][k]))
        text[n + i] = "(%s)" % ",".join(parts)
    return text[n + len(desc) - 1] + ";"


def reference_spr_sample(mods, desc, brlen, n, n_sample, rng):
    """n_sample SPR neighbours by the reference's own enumeration, spread over its branch list."""
    topo = mods["tree"].tree(arrays_to_newick(desc, brlen, n)).toTopology()
    branches = topo.getBranches()
    out_d, out_b = [], []
    order = rng.permutation(len(branches))
    per_branch = 2
    for bi in order:
        got = 0
        for mv in topo.iterSPRForBranch(branches[bi]):
            if rng.random() < 0.7:
                continue  # not always the first destinations
            d, b = newick_to_arrays(mv.toTree().getNewick(), n)
            out_d.append(d)
            out_b.append(b)
            got += 1
            if got >= per_branch or len(out_d) >= n_sample:
                break
        if len(out_d) >= n_sample:
            break
    return np.stack(out_d), np.stack(out_b)


def tbr_sample(desc, brlen, n, n_sample, rng):
    """Textbook TBR in plain Python on an unrooted adjacency list: cut an internal edge, re-root each
    part on one of its edges, join the two with a new edge (lengths: cut edge length for the new edge,
    split edges halved, merged edges summed)."""
    import networkx as nx
    out_d, out_b = [], []
    G0 = nx.Graph()
    root = n + len(desc) - 1
    for i, (a, b) in enumerate(desc):
        for k, c in enumerate((int(a), int(b))):
            G0.add_edge(n + i, c, w=float(brlen[i][k]))
    # the root of the descriptor is a degree-2 node of the unrooted tree: dissolve it
    (x, y) = list(G0.neighbors(root))
    G0.add_edge(x, y, w=G0[root][x]["w"] + G0[root][y]["w"])
    G0.remove_node(root)
    inner_edges = [(u, v) for u, v in G0.edges if u >= n and v >= n]
    nxt = [10 * n]

    def dissolve(G, v):
        a, b = list(G.neighbors(v))
        G.add_edge(a, b, w=G[v][a]["w"] + G[v][b]["w"])
        G.remove_node(v)

    def split(G, e):
        u, v = e
        w = G[u][v]["w"]
        G.remove_edge(u, v)
        m = nxt[0]
        nxt[0] += 1
        G.add_edge(u, m, w=w / 2)
        G.add_edge(m, v, w=w / 2)
        return m

    while len(out_d) < n_sample:
        G = G0.copy()
        u, v = inner_edges[rng.integers(len(inner_edges))]
        wcut = G[u][v]["w"]
        G.remove_edge(u, v)
        comp_u = nx.node_connected_component(G, u)
        dissolve(G, u)
        dissolve(G, v)
        eu = [e for e in G.edges if e[0] in comp_u]
        ev = [e for e in G.edges if e[0] not in comp_u]
        if not eu or not ev:
            continue
        mu = split(G, eu[rng.integers(len(eu))])
        mv = split(G, ev[rng.integers(len(ev))])
        G.add_edge(mu, mv, w=wcut)
        # root on tip 0's branch and write the descriptor in post-order
        d, b = [], []
        ids = {}

        def walk(x, parent):
            if x < n:
                return x
            kids = [c for c in G.neighbors(x) if c != parent]
            a = walk(kids[0], x)
            c = walk(kids[1], x)
            d.append((a, c))
            b.append((G[x][kids[0]]["w"], G[x][kids[1]]["w"]))
            ids[x] = n + len(d) - 1
            return ids[x]

        sys.setrecursionlimit(10000)
        w0 = list(G.neighbors(0))[0]
        top = walk(w0, 0)
        d.append((0, top))
        b.append((G[0][w0]["w"], 0.0))
        assert len(d) == n - 1
        out_d.append(np.array(d, dtype=np.int16))
        out_b.append(np.array(b))
    return np.stack(out_d), np.stack(out_b)


def main():
    if not refshim.available():
        sys.exit("needs /root/reference")
    stubs = {k: types.ModuleType(k) for k in ("alignment", "executable", "scoring")}
    mods = refshim.install(stubs, want=("base", "newick", "rearrangement", "tree"))
    rng = np.random.default_rng(20261020)
    out = {}
    for name, wl in bench.WORKLOADS.items():
        if wl["kind"] == "spr":
            d, br = bench.base_tree(wl, 0)
            sd, sb = reference_spr_sample(mods, d, br, wl["n_taxa"], wl["ref_sample"], rng)
        elif wl["kind"] == "tbr":
            d, br = bench.base_tree(wl, 0)
            sd, sb = tbr_sample(d, br, wl["n_taxa"], wl["ref_sample"], rng)
        else:
            continue
        out[name + "_desc"], out[name + "_brlen"] = sd, sb
        print(name, sd.shape, sb.shape)
    np.savez_compressed(os.path.join(HERE, "bench_ref_trees.npz"), **out)
    print("wrote", os.path.join(HERE, "bench_ref_trees.npz"), os.path.getsize(os.path.join(HERE, "bench_ref_trees.npz")), "bytes")


if __name__ == "__main__":
    main()
