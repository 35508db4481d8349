"""Deterministic synthetic workloads (trees, nucleotide/protein models, rate classes).

These are the inputs of bench.py and of the parity tests: numeric arrays in exactly the
form the C ABI (include/comap_b200.h) takes.  Shapes follow BASELINE.json's configs and
SURVEY.md s8(d) "Synthetic inputs".  Pure numpy/scipy; nothing here is on the hot path.
"""
import os
import numpy as np

AA_ORDER = "ARNDCQEGHILKMFPSTWYV"
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def random_tree(n_leaves, seed, mean_brlen, min_brlen=1e-6):
    """Random unrooted tree by sequential random joins, rooted at a trifurcation.

    Returns (parent[int32 n_nodes], brlen[float64 n_nodes]) with ids in Newick post-order
    (children before parent, root last, parent[root] = -1), n_nodes = 2*n_leaves - 2.
    """
    assert n_leaves >= 3
    rng = np.random.default_rng(seed)
    children = {}
    live = list(range(n_leaves))
    nxt = n_leaves
    while len(live) > 3:
        i, j = sorted(rng.choice(len(live), size=2, replace=False))
        a, b = live[i], live[j]
        children[nxt] = [a, b]
        live = [x for k, x in enumerate(live) if k not in (i, j)] + [nxt]
        nxt += 1
    children[nxt] = list(live)
    root = nxt
    # post-order relabel
    new_id = {}
    order = []
    stack = [(root, 0)]
    while stack:
        v, k = stack.pop()
        ch = children.get(v, [])
        if k < len(ch):
            stack.append((v, k + 1))
            stack.append((ch[k], 0))
        else:
            new_id[v] = len(order)
            order.append(v)
    n_nodes = len(order)
    parent = np.full(n_nodes, -1, dtype=np.int32)
    for v, ch in children.items():
        for c in ch:
            parent[new_id[c]] = new_id[v]
    brlen = np.maximum(rng.exponential(mean_brlen, size=n_nodes), min_brlen)
    brlen[n_nodes - 1] = 0.0
    return parent, brlen


def n_leaves_of(parent):
    has_child = np.zeros(len(parent), dtype=bool)
    has_child[parent[parent >= 0]] = True
    return int((~has_child).sum())


def hky85(kappa, pi):
    """HKY85 generator (states A C G T), normalised to one expected substitution per unit."""
    pi = np.asarray(pi, dtype=np.float64)
    pi = pi / pi.sum()
    Q = np.tile(pi, (4, 1))
    for a, b in ((0, 2), (2, 0), (1, 3), (3, 1)):
        Q[a, b] *= kappa
    return _finish(Q, pi)


def gtr(a, b, c, d, e, pi):
    """GTR with Bio++'s parameter naming: a=C<->T, b=A<->T, c=G<->T,
    d=A<->C, e=C<->G, A<->G fixed to 1; states A C G T."""
    pi = np.asarray(pi, dtype=np.float64)
    pi = pi / pi.sum()
    s = np.zeros((4, 4))
    s[0, 1] = s[1, 0] = d
    s[0, 2] = s[2, 0] = 1.0
    s[0, 3] = s[3, 0] = b
    s[1, 2] = s[2, 1] = e
    s[1, 3] = s[3, 1] = a
    s[2, 3] = s[3, 2] = c
    return _finish(s * pi[None, :], pi)


def jc(n_states):
    pi = np.full(n_states, 1.0 / n_states)
    return _finish(np.ones((n_states, n_states)) * pi[None, :], pi)


def read_paml_dat(path):
    rows = []
    for ln in open(path):
        ln = ln.strip()
        if not ln or ln.startswith("#"):
            continue
        rows.append([float(x) for x in ln.split()])
    n = len(rows)  # n-1 triangle rows + 1 frequency row
    S = np.zeros((n, n))
    for i in range(n - 1):
        assert len(rows[i]) == i + 1
        for j, v in enumerate(rows[i]):
            S[i + 1, j] = S[j, i + 1] = v
    pi = np.array(rows[-1])
    assert len(pi) == n
    return S, pi


def jtt92():
    S, pi = read_paml_dat(os.path.join(_DATA, "jtt92_dcmut.dat"))
    pi = pi / pi.sum()
    return _finish(S * pi[None, :], pi)


def _finish(Q, pi):
    Q = np.array(Q, dtype=np.float64)
    np.fill_diagonal(Q, 0.0)
    np.fill_diagonal(Q, -Q.sum(1))
    Q /= -(pi * np.diag(Q)).sum()
    return np.ascontiguousarray(Q), np.ascontiguousarray(pi)


def gamma_rates(alpha, n):
    """Bio++ Gamma(n, alpha): n equiprobable classes, class rate = conditional mean."""
    from scipy.stats import gamma as G
    q = G.ppf(np.arange(n + 1) / n, alpha, scale=1.0 / alpha)
    e = G.cdf(q, alpha + 1, scale=1.0 / alpha)
    return n * (e[1:] - e[:-1]), np.full(n, 1.0 / n)


def invariant(rates, probs, p):
    """Bio++ Invariant(dist=D, p=p): class 0 has rate 0 with probability p."""
    rates = np.concatenate([[0.0], np.asarray(rates) / (1.0 - p)])
    probs = np.concatenate([[p], np.asarray(probs) * (1.0 - p)])
    return rates, probs


def identity_code_mask(n_states):
    """Code table for simulated / fully resolved data: code k = state k, plus one 'unknown'."""
    m = [1 << k for k in range(n_states)] + [(1 << n_states) - 1]
    return np.array(m, dtype=np.uint32)
