"""BASELINE.json's two synthetic configurations at FULL size, through the C ABI.

The oracle cannot score 12.5 M pairs x 998 branches or cluster 20 000 sites in test time, so
parity at these sizes is shown through properties that do not depend on size:

  * sites are independent: a random SAMPLE of sites mapped by the oracle on the full 500 / 200
    taxon tree must equal the same rows of the device result (1e-9, north_star);
  * a permutation of the site order permutes the result (no dependence on chunk / lane / CTA);
  * the pair table restricted to a sample of sites equals the oracle's table of that sample,
    bit for bit, p-values against the FULL 10^6-sample null included;
  * replicate shards of the null concatenate to the one-call null (the multi-GPU partition);
  * single replicates of the null, re-simulated by the oracle at their global site indices
    (up to 2 * 10^6), give the same alignments and the same statistics;
  * dendrogram: first merge = global first minimum, complete-linkage heights non-decreasing
    and equal to the maximum distance between the two merged member sets, every id merged
    once, same multiset of heights as scipy's nn-chain; groups equal the oracle's walk of the
    device dendrogram."""
import numpy as np
import pytest
import oracle_binding as O
from comap_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture()
def ctx():
    from comap_b200 import api
    c = api.Context()
    yield c
    c.close()


def _check_rows(r, q, idx):
    assert np.allclose(r["n"][idx], q["n"], rtol=RTOL, atol=1e-14)
    assert np.allclose(r["norm"][idx], q["norm"], rtol=RTOL)
    assert np.allclose(r["loglik"][idx], q["loglik"], rtol=1e-12)
    assert np.allclose(r["post_rate"][idx], q["post_rate"], rtol=RTOL)
    assert np.array_equal(r["rate_class"][idx], q["rate_class"])


def _eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


def test_nucleotide_5000_sites_500_taxa_correlation_with_1000_replicates(ctx):
    """configs[3]: HKY85 + Gamma(4), all 12 497 500 pairs, 1000 x 1000 null samples."""
    S, T, rep_cpu, rep_ram, K = 5000, 500, 1000, 1000, 10
    parent, brlen = syn.random_tree(T, 20251018, 0.02)
    Q, pi = syn.hky85(2.5, [0.3, 0.2, 0.2, 0.3])
    rates, probs = syn.gamma_rates(0.5, 4)
    mask = syn.identity_code_mask(4)
    ctx.set_tree(parent, brlen); ctx.set_model(Q, pi, rates, probs)
    codes, _ = ctx.simulate(1, 0, S)
    assert np.array_equal(codes[:, :64], O.simulate(parent, brlen, Q, pi, rates, probs, 1, 0, 64)[0])
    rng = np.random.default_rng(7)

    # --- mapping: permutation, then sampled sites against the oracle on the full tree
    perm = rng.permutation(S)
    ctx.set_alignment(np.ascontiguousarray(codes[:, perm]), mask)
    rp = ctx.map()
    ctx.set_alignment(codes, mask)
    r = ctx.map()
    assert r["n"].shape == (S, len(parent) - 1)
    for key in ("n", "norm", "loglik", "post_rate", "rate_class"):
        assert np.array_equal(rp[key], r[key][perm]), key      # bit for bit: no dependence on position
    idx = np.sort(rng.choice(S, 96, replace=False))
    q = O.map_sites(parent, brlen, Q, pi, rates, probs, np.ascontiguousarray(codes[:, idx]), mask)
    _check_rows(r, q, idx)
    assert np.all(np.isfinite(r["n"])) and np.all(r["n"] >= 0) and np.all(r["loglik"] < 0)

    # --- null: one call, then 8 replicate shards (what 8 ranks compute)
    raw = ctx.null_intra("correlation", 2, rep_cpu, rep_ram, K=K, want_raw=True)
    null = ctx.null_get()
    assert raw.shape == (rep_cpu * rep_ram, 4)
    assert null["K"] == K and null["bin_offsets"][-1] <= rep_cpu * rep_ram
    assert abs(null["nmax"] - r["norm"].max()) <= 1e-12 * null["nmax"]
    for b in range(K):
        seg = null["sorted"][null["bin_offsets"][b]:null["bin_offsets"][b + 1]]
        seg = seg[~np.isnan(seg)]
        assert np.all(np.diff(seg) >= 0)                       # each bin sorted (CoETools.cpp:650-652)
    parts = [ctx.null_intra("correlation", 2, rep_cpu, rep_ram, K=0, rep_begin=125 * g, rep_end=125 * (g + 1),
                            want_raw=True) for g in range(8)]
    assert _eq(np.concatenate(parts), raw)
    # three replicates re-simulated and re-scored by the oracle (first 48 site pairs of each)
    for rep in (0, 517, 999):
        a = ctx.simulate(2, (2 * rep) * rep_ram, 48)[0]
        b = ctx.simulate(2, (2 * rep + 1) * rep_ram, 48)[0]
        assert np.array_equal(a, O.simulate(parent, brlen, Q, pi, rates, probs, 2, (2 * rep) * rep_ram, 48)[0])
        assert np.array_equal(b, O.simulate(parent, brlen, Q, pi, rates, probs, 2, (2 * rep + 1) * rep_ram, 48)[0])
        o = O.null_intra(parent, brlen, Q, pi, rates, probs, "correlation", a[None], b[None], K, null["nmax"])
        mine = raw[rep * rep_ram:rep * rep_ram + 48]
        fin = ~np.isnan(o["raw"][:, 0])
        assert np.array_equal(fin, ~np.isnan(mine[:, 0]))
        assert np.allclose(mine[fin, 0], o["raw"][fin, 0], rtol=RTOL, atol=1e-12)
        assert np.allclose(mine[:, 3], o["raw"][:, 3], rtol=RTOL)

    # --- the restored one-call null scores every pair
    ctx.null_intra("correlation", 2, rep_cpu, rep_ram, K=K)
    g, k = ctx.pairs("correlation", use_null=True)
    assert k == S * (S - 1) // 2
    key = g["i"].astype(np.int64) * S + g["j"]
    assert np.all(np.diff(key) > 0) and np.all(g["i"] < g["j"])   # reference row order, each pair once
    has = g["nsim"] > 0
    assert np.all(g["pvalue"][has] >= 1.0 / (g["nsim"][has] + 1.0)) and np.all(g["pvalue"][has] <= 1.0)
    assert np.all(np.isnan(g["pvalue"][~has]))
    fin = ~np.isnan(g["stat"])
    assert np.all(np.abs(g["stat"][fin]) <= 1 + 1e-12)
    # sample of sites: the oracle's table on the device vectors, p-values against the full null
    sub = np.sort(rng.choice(S, 260, replace=False))
    op = O.pairs("correlation", r["n"][sub], r["norm"][sub], r["post_rate"][sub], r["rate_class"][sub],
                 null=(K, null["nmax"], null["bin_offsets"], null["sorted"]))
    pos = np.full(S, -1); pos[sub] = np.arange(len(sub))
    sel = (pos[g["i"]] >= 0) & (pos[g["j"]] >= 0)
    assert sel.sum() == len(op["i"]) == 260 * 259 // 2
    assert np.array_equal(pos[g["i"][sel]], op["i"]) and np.array_equal(pos[g["j"][sel]], op["j"])
    assert _eq(g["stat"][sel], op["stat"])
    assert np.array_equal(g["nsim"][sel], op["nsim"]) and _eq(g["pvalue"][sel], op["pvalue"])
    assert np.array_equal(g["nmin"][sel], op["nmin"]) and np.array_equal(g["prmin"][sel], op["prmin"])
    assert np.array_equal(g["rcmin"][sel], op["rcmin"])


def _leaves(left, right, S, v):
    """Leaf set under node v (ids: leaves 0..S-1, merge k -> S + k)."""
    out, todo = [], [int(v)]
    while todo:
        u = todo.pop()
        if u < S:
            out.append(u)
        else:
            todo += [int(left[u - S]), int(right[u - S])]
    return np.array(out)


def _first_minimum(mat):
    """(i, j), i < j, of the first strictly-smallest off-diagonal entry in row-major order."""
    S = mat.shape[0]
    arg = np.array([int(np.argmin(mat[i, i + 1:])) + i + 1 for i in range(S - 1)])
    val = mat[np.arange(S - 1), arg]
    i = int(np.argmin(val))
    return i, int(arg[i])


def test_protein_20000_sites_200_taxa_clustering(ctx):
    """configs[4]: JTT92 + Gamma(4), correlation distance matrix (3.2 GB), complete linkage."""
    S, T = 20000, 200
    parent, brlen = syn.random_tree(T, 11, 0.05)
    Q, pi = syn.jtt92()
    rates, probs = syn.gamma_rates(0.8, 4)
    mask = syn.identity_code_mask(20)
    ctx.set_tree(parent, brlen); ctx.set_model(Q, pi, rates, probs)
    sim, _ = ctx.simulate(3, 0, S + S // 4)
    # input.remove_const + no duplicated columns (their vectors tie exactly), first S of them
    _, first = np.unique(sim, axis=1, return_index=True)
    first = np.sort(first)
    first = first[[len(set(sim[:, s])) > 1 for s in first]][:S]
    assert len(first) == S
    codes = np.ascontiguousarray(sim[:, first])
    ctx.set_alignment(codes, mask)
    r = ctx.map()
    rng = np.random.default_rng(8)
    idx = np.sort(rng.choice(S, 40, replace=False))
    q = O.map_sites(parent, brlen, Q, pi, rates, probs, np.ascontiguousarray(codes[:, idx]), mask)
    _check_rows(r, q, idx)

    mat = ctx.distance_matrix("correlation")
    sub = np.sort(rng.choice(S, 220, replace=False))
    assert np.array_equal(mat[np.ix_(sub, sub)], O.distance_matrix("correlation", r["n"][sub]))
    blk = mat[:3000, 17000:]
    assert np.array_equal(blk, mat[17000:, :3000].T) and np.all(np.diag(mat) == 0)

    left, right, height = ctx.cluster("complete")
    assert len(height) == S - 1 and np.all(np.diff(height) >= 0)
    used = np.concatenate([left, right])
    assert np.array_equal(np.sort(used), np.arange(2 * S - 2))      # every id merged exactly once
    # first merge: the first minimum in (i, j) order (the reference's tie-break)
    i0, j0 = _first_minimum(mat)
    assert {int(left[0]), int(right[0])} == {i0, j0} and 2 * height[0] == mat[i0, j0]
    for k in np.concatenate([rng.choice(S - 1, 60, replace=False), np.arange(S - 6, S - 1)]):
        a, b = _leaves(left, right, S, left[k]), _leaves(left, right, S, right[k])
        # height = len + (d / 2 - len), the reference's branch-length arithmetic: within an ulp of d / 2
        assert abs(2 * height[k] - mat[np.ix_(a, b)].max()) <= 4e-16 * 2 * height[k], k

    from scipy.cluster.hierarchy import linkage as sl
    from scipy.spatial.distance import squareform
    cond = squareform(mat, checks=False)
    Z = sl(cond, method="complete")
    del cond
    assert np.allclose(np.sort(2 * height), np.sort(Z[:, 2]), rtol=1e-12)

    g = ctx.groups("correlation", 10)
    og = O.groups("correlation", r["n"], r["norm"], left, right, height, 10)
    assert len(g["members"]) == len(og["members"]) > S // 20
    assert all(np.array_equal(a, b) for a, b in zip(g["members"], og["members"]))
    assert np.array_equal(g["height"], og["height"]) and np.array_equal(g["nmin"], og["nmin"])
    assert np.array_equal(g["stat"], og["stat"])
    assert max(len(m) for m in g["members"]) <= 10

    # clustering null: replicate shards concatenate
    both = ctx.cluster_null("correlation", "complete", seed=5, rep_begin=0, rep_end=2, max_size=10)
    one = ctx.cluster_null("correlation", "complete", seed=5, rep_begin=1, rep_end=2, max_size=10)
    sel = np.flatnonzero(both["rep"] == 1)
    assert len(sel) == len(one["rep"]) > 0
    assert np.array_equal(both["dmax"][sel], one["dmax"]) and np.array_equal(both["stat"][sel], one["stat"])
    assert all(np.array_equal(both["members"][k], m) for k, m in zip(sel, one["members"]))
