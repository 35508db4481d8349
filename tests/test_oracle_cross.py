"""Cross-checks of the C oracle against independent numpy / scipy restatements
(second opinions for the parts the reference ships no expected outputs for)."""
import numpy as np
import pytest
import helpers as H
import oracle_binding as O
from comap_b200 import synthetic as syn


@pytest.mark.parametrize("T,S,seed,amb", [(5, 40, 1, 0.0), (12, 64, 2, 0.1), (40, 50, 3, 0.02)])
def test_map_vs_numpy_dna(T, S, seed, amb):
    c = H.random_dna_case(T, S, seed, ambiguity=amb)
    r = O.map_sites(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], c["codes"], c["code_mask"])
    q = H.map_np(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], c["codes"], c["code_mask"])
    assert np.allclose(r["n"], q["n"], rtol=1e-9, atol=1e-14)
    assert np.allclose(r["norm"], q["norm"], rtol=1e-10)
    assert np.allclose(r["loglik"], q["loglik"], rtol=1e-11)
    assert np.allclose(r["post_rate"], q["post_rate"], rtol=1e-10)
    assert np.array_equal(r["rate_class"], q["rate_class"])


def test_map_invariant_class_and_gtr():
    parent, brlen = syn.random_tree(9, 7, 0.1)
    Q, pi = syn.gtr(1.6, 0.55, 0.35, 0.30, 0.28, [0.25, 0.2, 0.3, 0.25])
    rates, probs = syn.invariant(*syn.gamma_rates(0.737, 4), p=0.3666)
    assert abs((rates * probs).sum() - 1) < 1e-12 and rates[0] == 0
    rng = np.random.default_rng(5)
    codes = H.simulate_np(parent, brlen, Q, pi, rates, rng, 80)
    mask = syn.identity_code_mask(4)
    r = O.map_sites(parent, brlen, Q, pi, rates, probs, codes, mask)
    q = H.map_np(parent, brlen, Q, pi, rates, probs, codes, mask)
    assert np.all(np.isfinite(r["n"]))
    assert np.allclose(r["n"], q["n"], rtol=1e-9, atol=1e-14)
    assert np.array_equal(r["rate_class"], q["rate_class"])


def test_mapping_total_equals_brute_force_tiny_tree():
    """3 taxa, 1 class: posterior expected counts by explicit enumeration of ancestral states."""
    parent = np.array([3, 3, 3, -1], np.int32)
    brlen = np.array([0.3, 0.1, 0.7, 0.0])
    Q, pi = syn.hky85(3.0, [0.1, 0.2, 0.3, 0.4])
    rates, probs = np.array([1.0]), np.array([1.0])
    codes = np.array([[0, 1, 2], [0, 3, 2], [1, 3, 2]], np.uint8)
    r = O.map_sites(parent, brlen, Q, pi, rates, probs, codes, syn.identity_code_mask(4))
    P = [O.pmatrix(Q, pi, t) for t in brlen[:3]]
    N = [O.counts(Q, pi, t) for t in brlen[:3]]
    for s in range(3):
        tip = codes[:, s]
        w = np.array([pi[x] * P[0][x, tip[0]] * P[1][x, tip[1]] * P[2][x, tip[2]] for x in range(4)])
        for b in range(3):
            e = sum(w[x] * N[b][x, tip[b]] for x in range(4)) / w.sum()
            assert abs(r["n"][s, b] - e) < 1e-13
        assert abs(r["loglik"][s] - np.log(w.sum())) < 1e-13


def test_pmatrix_and_counts_vs_numpy():
    Q, pi = syn.jtt92()
    for t in (1e-7, 0.01, 0.5, 3.0):
        # eigen-based P(t) carries ~1e-16 absolute noise per entry (both sides)
        assert np.allclose(O.pmatrix(Q, pi, t), H.expm_rev(Q, pi, t), rtol=1e-10, atol=1e-14)
        Pn = H.expm_rev(Q, pi, t)
        assert np.allclose(O.counts(Q, pi, t) * O.pmatrix(Q, pi, t), H.unif_counts_np(Q, pi, t) * Pn, rtol=1e-10, atol=1e-15)
    w = np.abs(np.subtract.outer(np.arange(20.0), np.arange(20.0)))
    assert np.allclose(O.counts(Q, pi, 0.3, weights=w), H.unif_counts_np(Q, pi, 0.3, w), rtol=1e-10)
    assert np.all(O.counts(Q, pi, 0.0) == 0)


def test_statistics_vs_numpy():
    rng = np.random.default_rng(0)
    a, b = rng.gamma(0.3, 1.0, 97), rng.gamma(0.3, 1.0, 97)
    assert abs(O.stat("correlation", a, b) - np.corrcoef(a, b)[0, 1]) < 1e-13
    assert abs(O.stat("covariance", a, b) - np.cov(a, b)[0, 1]) < 1e-13
    assert abs(O.stat("cosinus", a, b) - a @ b / np.linalg.norm(a) / np.linalg.norm(b)) < 1e-13
    assert O.stat("cosubstitution", a, b) == ((a >= 1) & (b >= 1)).sum()
    comp = 1 - np.linalg.norm(a + b) / (np.linalg.norm(a) + np.linalg.norm(b))
    assert abs(O.stat("compensation", a, -b) - (1 - np.linalg.norm(a - b) / (np.linalg.norm(a) + np.linalg.norm(b)))) < 1e-13
    assert abs(O.stat("compensation", a, b) - comp) < 1e-13
    assert np.isnan(O.stat("correlation", np.ones(5), a[:5]))
    n = rng.gamma(0.3, 1.0, (6, 30))
    g = O.stat_group("correlation", n, [0, 2, 5])
    assert abs(g - min(np.corrcoef(n[[0, 2, 5]])[np.triu_indices(3, 1)])) < 1e-13
    gc = O.stat_group("compensation", n, [1, 2, 3, 4])
    assert abs(gc - (1 - np.linalg.norm(n[1:5].sum(0)) / np.linalg.norm(n[1:5], axis=1).sum())) < 1e-13


def test_domain_index_semantics():
    # Domain.cpp:113-122: [lo, hi) with equal-width bins; upper bound excluded
    assert O.domain_index(0, 10, 10, 0.0) == 0
    assert O.domain_index(0, 10, 10, 9.999) == 9
    assert O.domain_index(0, 10, 10, 10.0) == -1
    assert O.domain_index(0, 10, 10, -1e-9) == -1
    assert O.domain_index(0, 1, 4, 0.25) == 1


def test_pairs_filters_and_pvalues():
    rng = np.random.default_rng(3)
    S, B = 30, 41
    n = rng.gamma(0.4, 1.0, (S, B)); norm = np.sqrt((n ** 2).sum(1))
    pr = rng.gamma(2, 0.5, S); rc = rng.integers(0, 4, S).astype(np.int32)
    K, nmax = 5, norm.max()
    null_stat = rng.uniform(-1, 1, 4000); null_n = rng.uniform(0, nmax * 1.1, 4000)
    bins = [np.sort(null_stat[[O.domain_index(0, nmax, K, x) == k for x in null_n]]) for k in range(K)]
    offs = np.concatenate([[0], np.cumsum([len(b) for b in bins])])
    r = O.pairs("correlation", n, norm, pr, rc, null=(K, nmax, offs, np.concatenate(bins)))
    assert len(r["i"]) == S * (S - 1) // 2
    assert np.all(r["i"][1:] * S + r["j"][1:] > r["i"][:-1] * S + r["j"][:-1])  # reference order
    cc = np.corrcoef(n)
    assert np.allclose(r["stat"], cc[r["i"], r["j"]], atol=1e-13)
    k = 17
    i, j = r["i"][k], r["j"][k]
    nm = min(norm[i], norm[j]); cat = O.domain_index(0, nmax, K, nm)
    cnt = (bins[cat] < r["stat"][k]).sum()
    assert r["nsim"][k] == len(bins[cat]) and r["pvalue"][k] == (len(bins[cat]) - cnt + 1) / (len(bins[cat]) + 1)
    # the pair holding the max norm falls out of [0, max) -> NA / 0  (CoETools.cpp:718-720)
    top = np.argmax(norm)
    sel = (np.minimum(norm[r["i"]], norm[r["j"]]) == norm[top])
    assert np.all(np.isnan(r["pvalue"][sel])) and np.all(r["nsim"][sel] == 0)
    f = O.pairs("correlation", n, norm, pr, rc, min_rate_class=1, min_rate=0.5, max_rate_class_diff=1,
                max_rate_diff=0.8, min_stat=0.1)
    ok = (rc[r["i"]] >= 1) & (rc[r["j"]] >= 1) & (pr[r["i"]] >= 0.5) & (pr[r["j"]] >= 0.5) \
        & (np.abs(rc[r["i"]] - rc[r["j"]]) <= 1) & (np.abs(pr[r["i"]] - pr[r["j"]]) <= 0.8) & (np.abs(r["stat"]) >= 0.1)
    assert np.array_equal(f["i"], r["i"][ok]) and np.array_equal(f["j"], r["j"][ok])


def test_simulate_is_counter_based_and_plausible():
    parent, brlen = syn.random_tree(8, 11, 0.2)
    Q, pi = syn.hky85(2.0, [0.4, 0.1, 0.2, 0.3])
    rates, probs = syn.gamma_rates(0.7, 4)
    a, ca = O.simulate(parent, brlen, Q, pi, rates, probs, 42, 0, 4000)
    b, cb = O.simulate(parent, brlen, Q, pi, rates, probs, 42, 1000, 500)
    assert np.array_equal(a[:, 1000:1500], b) and np.array_equal(ca[1000:1500], cb)
    c, _ = O.simulate(parent, brlen, Q, pi, rates, probs, 43, 0, 4000)
    assert not np.array_equal(a, c)
    freq = np.bincount(a.ravel(), minlength=4) / a.size
    assert np.allclose(freq, pi, atol=0.03)
    assert np.allclose(np.bincount(ca, minlength=4) / 4000, 0.25, atol=0.03)


def test_null_intra_bins_and_raw():
    parent, brlen = syn.random_tree(7, 5, 0.15)
    Q, pi = syn.hky85(2.0, [0.25] * 4)
    rates, probs = syn.gamma_rates(0.5, 4)
    rep_cpu, rep_ram = 3, 50
    T = 7
    s1 = np.stack([O.simulate(parent, brlen, Q, pi, rates, probs, 9, (2 * i) * rep_ram, rep_ram)[0] for i in range(rep_cpu)])
    s2 = np.stack([O.simulate(parent, brlen, Q, pi, rates, probs, 9, (2 * i + 1) * rep_ram, rep_ram)[0] for i in range(rep_cpu)])
    assert s1.shape == (rep_cpu, T, rep_ram)
    K, nmax = 4, 1.5
    r = O.null_intra(parent, brlen, Q, pi, rates, probs, "correlation", s1, s2, K, nmax)
    mask = syn.identity_code_mask(4)
    m1 = O.map_sites(parent, brlen, Q, pi, rates, probs, s1[1], mask)
    m2 = O.map_sites(parent, brlen, Q, pi, rates, probs, s2[1], mask)
    j = 7; k = rep_ram + j
    st = O.stat("correlation", m1["n"][j], m2["n"][j])
    assert r["raw"][k, 0] == st or (np.isnan(st) and np.isnan(r["raw"][k, 0]))
    assert r["raw"][k, 3] == min(m1["norm"][j], m2["norm"][j])
    assert r["raw"][k, 1] == min(m1["rate_class"][j], m2["rate_class"][j])
    cat = np.array([O.domain_index(0, nmax, K, x) for x in r["raw"][:, 3]])
    for b in range(K):
        seg = r["sorted"][r["bin_offsets"][b]:r["bin_offsets"][b + 1]]
        assert len(seg) == (cat == b).sum()
        fin = seg[~np.isnan(seg)]
        assert np.all(np.diff(fin) >= 0)


@pytest.mark.parametrize("linkage,method", [("complete", "complete"), ("single", "single"), ("average", "average")])
def test_hclust_vs_scipy(linkage, method):
    from scipy.cluster.hierarchy import linkage as sl
    from scipy.spatial.distance import squareform
    rng = np.random.default_rng(1)
    n = rng.gamma(0.4, 1.0, (25, 60))
    mat = O.distance_matrix("correlation", n)
    assert np.allclose(mat, 1 - np.corrcoef(n), atol=1e-13) and np.all(np.diag(mat) == 0)
    left, right, height = O.hclust(linkage, mat)
    Z = sl(squareform(mat, checks=False), method=method)
    assert np.allclose(np.sort(2 * height), np.sort(Z[:, 2]), rtol=1e-12)
    g = O.groups("correlation", n, np.sqrt((n ** 2).sum(1)), left, right, height, max_size=10)
    assert all(len(m) <= 10 for m in g["members"])
    assert np.allclose(g["stat"], 1 - 2 * g["height"])
    norm = np.sqrt((n ** 2).sum(1))
    assert np.allclose(g["nmin"], [norm[m].min() for m in g["members"]])
    if linkage == "complete":
        # complete linkage: Dmax = max pairwise distance inside the group
        for m, h in zip(g["members"], g["height"]):
            assert abs(2 * h - mat[np.ix_(m, m)].max()) < 1e-12


def test_groups_compensation_and_euclid():
    rng = np.random.default_rng(2)
    n = rng.normal(0, 1, (12, 30)); norm = np.sqrt((n ** 2).sum(1))
    mat = O.distance_matrix("compensation", n)
    i, j = 3, 8
    assert abs(mat[i, j] - np.linalg.norm(n[i] + n[j]) / (norm[i] + norm[j])) < 1e-13
    left, right, height = O.hclust("complete", mat)
    g = O.groups("compensation", n, norm, left, right, height, max_size=12)
    assert len(g["members"]) == 11 and sorted(g["members"][-1]) == list(range(12))
    for m, s in zip(g["members"], g["stat"]):
        assert abs(s - (1 - np.linalg.norm(n[m].sum(0)) / norm[m].sum())) < 1e-13
    me = O.distance_matrix("euclidian", n)
    assert abs(me[i, j] - np.linalg.norm(n[i] - n[j])) < 1e-13


def test_continuous_rate_simulator_distribution():
    """simulations.continuous = yes (oracle side): site rates follow Gamma(alpha, beta = alpha) -- mean 1,
    variance 1 / alpha, and the invariant mixture keeps mean 1 with a mass p at 0; faster sites differ more
    from the root state than slow ones."""
    from scipy import stats
    parent, brlen = syn.random_tree(12, 3, 0.2)
    Q, pi = syn.hky85(2.0, [0.3, 0.2, 0.2, 0.3])
    for alpha in (0.5, 2.0):
        st, r = O.simulate_continuous(parent, brlen, Q, pi, "gamma", alpha, 0.0, 11, 0, 20000)
        assert abs(r.mean() - 1.0) < 0.03 and abs(r.var() - 1.0 / alpha) < 0.12 / alpha
        assert stats.kstest(r, "gamma", args=(alpha, 0, 1.0 / alpha)).pvalue > 1e-3
        var_sites = (st != st[0]).any(axis=0)
        assert var_sites[r > np.median(r)].mean() > var_sites[r <= np.median(r)].mean() + 0.1
    st, r = O.simulate_continuous(parent, brlen, Q, pi, "invariant", 1.0, 0.3, 5, 0, 20000)
    assert abs((r == 0).mean() - 0.3) < 0.02 and abs(r.mean() - 1.0) < 0.04
    assert not (st[:, r == 0] != st[0, r == 0]).any()          # invariant sites never change
    st, r = O.simulate_continuous(parent, brlen, Q, pi, "constant", 1.0, 0.0, 5, 0, 100)
    assert np.all(r == 1.0)


def test_mapping_variants_equal_brute_force_enumeration():
    """nijt.average / nijt.joint (CoETools.cpp:393-407): the three other mapping functions as restated in the oracle
    ([Bio++ / from memory]) against an explicit enumeration of every ancestral-state assignment of a 5-taxon tree
    with two rate classes -- joint and marginal posteriors by summation, not by recursion."""
    import itertools
    #        leaves 0 1 | inner 2 = (0,1) | leaves 3 4 | inner 5 = (3,4) | leaf 6 | root 7 = (2,5,6)
    parent = np.array([2, 2, 7, 5, 5, 7, 7, -1], np.int32)
    brlen = np.array([0.3, 0.15, 0.2, 0.4, 0.05, 0.6, 0.25, 0.0])
    Q, pi = syn.hky85(3.0, [0.1, 0.2, 0.3, 0.4])
    rates, probs = np.array([0.3, 1.7]), np.array([0.5, 0.5])
    codes = np.array([[0, 1, 2, 3], [0, 3, 2, 1], [1, 3, 2, 4], [1, 0, 2, 2], [2, 0, 3, 2]], np.uint8)  # code 4 = unknown
    mask = np.array([1, 2, 4, 8, 15], np.uint32)
    leaves, inner = [0, 1, 3, 4, 6], [2, 5, 7]
    B, S, A, C = 7, codes.shape[1], 4, 2
    P = np.array([[O.pmatrix(Q, pi, brlen[v] * r) for r in rates] for v in range(B)])
    N = np.array([[O.counts(Q, pi, brlen[v] * r) for r in rates] for v in range(B)])
    tipvec = ((mask[codes][:, :, None] >> np.arange(A)) & 1).astype(float)            # [leaf row][site][state]
    # weight of (class, states of the inner nodes) given the data, per site
    Wt = np.zeros((S, C, A, A, A))
    for s, c, x2, x5, x7 in itertools.product(range(S), range(C), range(A), range(A), range(A)):
        st = {2: x2, 5: x5, 7: x7}
        w = probs[c] * pi[x7]
        for v in range(B):
            f = st[parent[v]]
            w *= P[v, c, f, st[v]] if v in st else P[v, c, f] @ tipvec[leaves.index(v), s]
        Wt[s, c, x2, x5, x7] = w
    L = Wt.sum(axis=(1, 2, 3, 4))
    ax = {2: 2, 5: 3, 7: 4}

    def marg(s, v):  # [class][state] weight of inner node v
        other = tuple(a for n_, a in ax.items() if n_ != v)
        return Wt[s].sum(axis=tuple(o - 1 for o in other))

    exp = {k: np.zeros((S, B)) for k in ((1, 0), (0, 1), (0, 0))}
    for s in range(S):
        anc = {v: int(np.argmax((marg(s, v) / L[s]).sum(0))) for v in inner}
        for v in range(B):
            f = parent[v]
            mf = marg(s, f)
            if v in ax:  # joint weights of (class, x at f, y at v)
                keep = [0, ax[f] - 1, ax[v] - 1]
                J = Wt[s].sum(axis=tuple(i for i in range(4) if i not in keep))
                if ax[f] > ax[v]: J = J.transpose(0, 2, 1)
                mv = marg(s, v) / marg(s, v).sum()
                yv = anc[v]
            else:        # leaf: joint weight with each compatible tip state
                t = tipvec[leaves.index(v), s]
                up = np.array([[mf[c, x] / max(P[v, c, x] @ t, 1e-300) for x in range(A)] for c in range(C)])
                J = up[:, :, None] * P[v] * t[None, None, :]
                mv = t[None, :] * probs[:, None]
                yv = int(np.argmax(t))
            pf = mf / mf.sum()
            exp[(1, 0)][s, v] = sum(pf[c, x] * mv[c, y] * N[v, c, x, y] for c in range(C) for x in range(A) for y in range(A))
            Jc = J.sum(0)
            bx, by = np.unravel_index(np.argmax(Jc), Jc.shape)
            exp[(0, 1)][s, v] = (J[:, bx, by] * N[v, :, bx, by]).sum() / Jc[bx, by]
            exp[(0, 0)][s, v] = (N[v, :, anc[f], yv] * probs).sum()
    try:
        for (av, jo), e in exp.items():
            O.set_map_mode(av, jo)
            r = O.map_sites(parent, brlen, Q, pi, rates, probs, codes, mask)
            assert np.allclose(r["n"], e, rtol=1e-10, atol=1e-13), (av, jo, np.abs(r["n"] - e).max())
            assert np.allclose(r["loglik"], np.log(L), rtol=1e-12)
    finally:
        O.set_map_mode()


def test_label_count_and_label_mutual_information():
    """nijt=Label (one label per substitution type) and statistic=MI over those labels (CoETools.cpp:577-589):
    numpy restatement of the joint-table formula on random label vectors."""
    Q, pi = syn.hky85(2.0, [0.25, 0.25, 0.25, 0.25])
    Nl = O.counts(Q, pi, 0.3, method="label")
    assert np.array_equal(Nl, np.array([[0, 1, 2, 3], [4, 0, 5, 6], [7, 8, 0, 9], [10, 11, 12, 0]], float))
    rng = np.random.default_rng(5)
    O.set_mi_label(4)
    for B in (7, 60, 400):
        a = rng.choice(13, size=B, p=[0.7] + [0.025] * 12).astype(float)
        b = np.where(rng.random(B) < 0.5, a, rng.choice(13, size=B)).astype(float)
        e = 0.0
        for u in np.unique(a):
            for w in np.unique(b):
                n12 = np.sum((a == u) & (b == w))
                if n12:
                    e += n12 / B * np.log(n12 * B / (np.sum(a == u) * np.sum(b == w))) / np.log(2.7182818)
        assert abs(O.stat("mi_label", a, b) - e) < 1e-12
        assert abs(O.stat("mi_label", a + 1e-13, b - 1e-13) - e) < 1e-12   # labels recovered through the +-0.5 bounds


def test_prob_one_jump_count():
    """nijt=ProbOneJump: 1 when the ends differ; for equal ends 1 - P(no event on the branch) / P_xx(t), which the
    converged uniformization series confirms: P(no event | x -> x) = exp(Q_xx t) / P_xx(t)."""
    Q, pi = syn.hky85(2.5, [0.3, 0.2, 0.2, 0.3])
    for t in (1e-6, 0.05, 0.7, 3.0):
        N = O.counts(Q, pi, t, method="one_jump")
        P = O.pmatrix(Q, pi, t)
        off = ~np.eye(4, dtype=bool)
        assert np.all(N[off] == 1.0)
        assert np.allclose(np.diag(N), 1 - np.exp(np.diag(Q) * t) / np.diag(P), rtol=1e-12, atol=1e-15)
        assert np.all((np.diag(N) >= -1e-15) & (np.diag(N) < 1))
    # more substitutions are expected than the probability of at least one, and they agree on short branches
    Nu = O.counts(Q, pi, 1e-3)
    N1 = O.counts(Q, pi, 1e-3, method="one_jump")
    assert np.all(Nu >= N1 - 1e-12) and np.allclose(Nu[off], N1[off], atol=2e-3)


def _np_pair(c1, c2, A, mask):
    """numpy restatement of mutualInformation / jointEntropy with unknowns resolved (joint table by outer products)."""
    P = np.zeros((A, A))
    for a, b in zip(c1, c2):
        v1 = ((mask[a] >> np.arange(A)) & 1).astype(float); v2 = ((mask[b] >> np.arange(A)) & 1).astype(float)
        if v1.sum() and v2.sum():
            P += np.outer(v1 / v1.sum(), v2 / v2.sum())
    P /= P.sum()
    p1, p2 = P.sum(1), P.sum(0)
    nz = P > 0
    return float((P[nz] * np.log(P[nz] / np.outer(p1, p2)[nz])).sum()), float(-(P[nz] * np.log(P[nz])).sum())


def test_mica_site_statistics_vs_numpy():
    """Mica (CoMap/Mica.cpp:92,341-361): entropy, MI and joint entropy of alignment columns, unknown characters resolved
    over their compatible states."""
    rng = np.random.default_rng(3)
    mask = np.array([1, 2, 4, 8, 15, 5, 0], np.uint32)       # A C G T N R(A|G) and a character outside the alphabet
    codes = rng.choice(7, size=(40, 12), p=[0.3, 0.25, 0.2, 0.15, 0.05, 0.04, 0.01]).astype(np.uint8)
    codes[:, 3] = codes[:, 2]                                  # identical columns: MI = H = Hjoint
    codes[:, 5] = 1                                            # constant column: MI = 0 with everything
    for i, j in ((0, 1), (2, 3), (5, 7), (4, 9), (11, 0)):
        mi, hj = O.site_pair(codes[:, i], codes[:, j], 4, mask)
        emi, ehj = _np_pair(codes[:, i], codes[:, j], 4, mask)
        assert abs(mi - emi) < 1e-12 and abs(hj - ehj) < 1e-12
    mi, hj = O.site_pair(codes[:, 5], codes[:, 7], 4, mask)
    assert abs(mi) < 1e-15
    res = codes[:, 2] < 4                                      # entropy of the resolved part by hand
    col = np.where(codes[:, 2] < 4, codes[:, 2], 0)
    h = O.site_entropy(col, 4, mask)
    f = np.bincount(col, minlength=4) / len(col)
    assert abs(h + (f[f > 0] * np.log(f[f > 0])).sum()) < 1e-13
    H, avg = O.mica_sites(codes, 4, mask)
    assert abs(H[2] - O.site_entropy(codes[:, 2], 4, mask)) == 0
    e = np.mean([O.site_pair(codes[:, 4], codes[:, j], 4, mask)[0] for j in range(12) if j != 4])
    assert abs(avg[4] - e) < 1e-13


def _philox(c, k0, k1):
    c = list(c)
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xffffffff, p1 & 0xffffffff, ((p0 >> 32) ^ c[3] ^ k1) & 0xffffffff, p0 & 0xffffffff]
        k0 = (k0 + 0x9E3779B9) & 0xffffffff; k1 = (k1 + 0xBB67AE85) & 0xffffffff
    return c


def _py_mi_test(c1, c2, A, mask, seed, pair, max_perm):
    """miTest (Mica.cpp:92-118) in plain Python: the stopping rule and the shuffle stream the device and the oracle share
    (inside-out Fisher-Yates of the original columns, see orc_mica_permutation_test)."""
    full = (1 << A) - 1
    def constant(c):
        seen = {int(x) for x in c if (int(mask[x]) & full) != full}
        return len(seen) <= 1
    mi = O.site_pair(c1, c2, A, mask)[0]
    if constant(c1) or constant(c2):
        return 1.0, 0
    T = len(c1)
    count = i = 0
    while count < 5 and i < max_perm:
        s = []
        for column, src in enumerate((c1, c2)):                # every shuffle starts from the original column
            col = [int(src[0])] + [0] * (T - 1)
            k, blk = 1, 0
            while k < T:
                for w in _philox([pair & 0xffffffff, pair >> 32, i, (column << 24) | blk], seed & 0xffffffff, seed >> 32):
                    for draw in range(2 if T <= 256 else 1):       # a word serves two small ranges
                        if k < T:
                            pr = w * (k + 1)
                            p, w = pr >> 32, pr & 0xffffffff
                            col[k] = col[p]; col[p] = int(src[k])
                            k += 1
                blk += 1
            assert sorted(col) == sorted(int(x) for x in src)
            s.append(col)
        rep = O.site_pair(np.array(s[0], np.uint8), np.array(s[1], np.uint8), A, mask)[0]
        count += rep >= mi
        i += 1
    return (count + 1) / (i + 1), i


def test_mica_permutation_test_vs_python():
    """null.method = permutations: the oracle's miTest against a plain-Python restatement of the same stopping rule,
    constant-column rule and Philox-driven Fisher-Yates shuffles; the shuffles keep the column's composition."""
    rng = np.random.default_rng(8)
    mask = np.array([1, 2, 4, 8, 15, 5], np.uint32)
    codes = rng.choice(6, size=(18, 9), p=[0.3, 0.25, 0.2, 0.15, 0.06, 0.04]).astype(np.uint8)
    codes[:, 4] = 3; codes[2, 4] = 4                           # constant but for an unknown character
    codes[:, 6] = codes[:, 5]                                  # a perfectly coupled pair: needs its whole budget
    pv, nb, closest = O.mica_permutations(codes, 4, mask, 99, 40)
    ii, jj = np.triu_indices(9, 1)
    for t in range(len(ii)):
        epv, enb = _py_mi_test(codes[:, ii[t]], codes[:, jj[t]], 4, mask, 99, t, 40)
        assert pv[t] == epv and nb[t] == enb, (t, pv[t], epv, nb[t], enb)
    const = (ii == 4) | (jj == 4)
    assert np.all(nb[const] == 0) and np.all(pv[const] == 1.0) and np.all(nb[~const] >= 5)
    t56 = int(np.where((ii == 5) & (jj == 6))[0][0])
    assert nb[t56] == 40 and pv[t56] < 0.2
    assert np.all(pv[~const] == np.where(nb[~const] < 40, 6.0 / (nb[~const] + 1.0), pv[~const]))


def test_mica_shuffle_stream_is_uniform():
    """The inside-out Fisher-Yates pass on the shared Philox stream (two draws per word) visits the 24 arrangements of
    four distinct characters evenly: chi-square over 12 000 shuffles stays far below the 0.1 % point (49.7, 23 d.o.f.)."""
    import collections
    src, T = [0, 1, 2, 3], 4
    seen = collections.Counter()
    n = 12000
    for i in range(n):
        col = [src[0], 0, 0, 0]
        k = 1
        for w in _philox([5, 0, i, 0], 77, 1):
            for draw in range(2):
                if k < T:
                    pr = w * (k + 1)
                    p, w = pr >> 32, pr & 0xffffffff
                    col[k] = col[p]; col[p] = src[k]
                    k += 1
        seen[tuple(col)] += 1
    assert len(seen) == 24
    chi2 = sum((v - n / 24) ** 2 / (n / 24) for v in seen.values())
    assert chi2 < 49.7, chi2
