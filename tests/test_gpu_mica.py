"""GPU parity of Mica's path (CoMap/Mica.cpp): column statistics, parametric-bootstrap null and the p-value epilogue,
against the CPU oracle through the C ABI.  MI / entropies at 1e-9 (device and host log differ in the last bit)."""
import numpy as np
import pytest
import helpers as H
import oracle_binding as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from comap_b200 import api
    c = api.Context()
    yield c
    c.close()


def _setup(ctx, c):
    ctx.set_tree(c["parent"], c["brlen"])
    ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
    ctx.set_alignment(c["codes"], c["code_mask"])


@pytest.mark.parametrize("case", ["dna", "protein"])
def test_site_statistics_and_pairs_vs_oracle(ctx, case):
    c = H.random_dna_case(30, 90, 5, mean_brlen=0.15, ambiguity=0.06) if case == "dna" else H.myoglobin_inputs()
    A = len(c["pi"])
    _setup(ctx, c)
    r = ctx.map()
    h, avg = ctx.mica_sites()
    oh, oavg = O.mica_sites(c["codes"], A, c["code_mask"])
    assert np.allclose(h, oh, rtol=1e-9, atol=1e-13) and np.allclose(avg, oavg, rtol=1e-9, atol=1e-13)
    p = ctx.mica_pairs("nmin")
    S = c["codes"].shape[1]
    ii, jj = np.triu_indices(S, 1)
    assert np.array_equal(p["i"], ii) and np.array_equal(p["j"], jj)           # mica's order: i ascending, j ascending
    sel = np.random.default_rng(1).choice(len(ii), size=min(len(ii), 600), replace=False)
    for t in sel:
        mi, hj = O.site_pair(c["codes"][:, ii[t]], c["codes"][:, jj[t]], A, c["code_mask"])
        assert abs(p["mi"][t] - mi) <= 1e-9 * max(1.0, abs(mi)) and abs(p["hjoint"][t] - hj) <= 1e-9 * max(1.0, abs(hj))
    assert np.array_equal(p["hmin"], np.minimum(h[ii], h[jj]))
    assert np.array_equal(p["nmin"], np.minimum(r["norm"][ii], r["norm"][jj]))
    assert p["mi"].min() > -1e-12 and np.all(p["mi"] <= p["hmin"] + 1e-9)        # 0 <= MI <= min entropy
    # listed pairs (the nonparametric bootstrap's resampled sites), in both orders
    a = np.array([0, 5, 7, 7, S - 1]); b = np.array([3, 5, 2, S - 1, 0])
    mi, hj = ctx.mica_pair_list(a, b)
    for k in range(len(a)):
        emi, ehj = O.site_pair(c["codes"][:, a[k]], c["codes"][:, b[k]], A, c["code_mask"])
        assert abs(mi[k] - emi) < 1e-9 and abs(hj[k] - ehj) < 1e-9


def test_parametric_bootstrap_null_and_pvalues(ctx):
    """null.method = parametric-bootstrap (Mica.cpp:470-545): the simulated pairs re-scored by the oracle from the
    exported alignments, norms against the oracle's mapping, and the p-values of the table against the device's own
    sorted null (Mica.cpp:672-683)."""
    c = H.random_dna_case(20, 70, 9, mean_brlen=0.12)
    _setup(ctx, c)
    r = ctx.map()
    rep_cpu, rep_ram, K = 3, 150, 4
    raw = ctx.mica_null_parametric(17, rep_cpu, rep_ram, K=K)
    for rep in range(rep_cpu):
        s1 = ctx.simulate(17, (2 * rep) * rep_ram, rep_ram)[0]
        s2 = ctx.simulate(17, (2 * rep + 1) * rep_ram, rep_ram)[0]
        m1 = O.map_sites(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], s1, c["code_mask"])
        m2 = O.map_sites(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], s2, c["code_mask"])
        rows = raw[rep * rep_ram:(rep + 1) * rep_ram]
        assert np.allclose(rows[:, 2], np.minimum(m1["norm"], m2["norm"]), rtol=1e-9, atol=1e-12)
        for j in range(0, rep_ram, 7):
            mi, hj = O.site_pair(s1[:, j], s2[:, j], 4, c["code_mask"])
            assert abs(rows[j, 0] - mi) < 1e-9 and abs(rows[j, 1] - hj) < 1e-9
    g = ctx.null_get()
    assert g["K"] == K and abs(g["nmax"] - r["norm"].max()) <= 1e-12 * g["nmax"]
    p = ctx.mica_pairs("nmin", use_null=True)
    off, srt = g["bin_offsets"], g["sorted"]
    for t in range(len(p["i"])):
        cat = O.domain_index(0.0, g["nmax"], K, p["nmin"][t])
        if cat < 0:
            assert np.isnan(p["pvalue"][t]) and p["nsim"][t] == 0
            continue
        b = srt[off[cat]:off[cat + 1]]
        cnt = int(np.searchsorted(b, p["mi"][t], side="left"))
        assert p["nsim"][t] == len(b) and p["pvalue"][t] == (len(b) - cnt + 1) / (len(b) + 1)
    # a null built on the host (z-score / nonparametric bootstrap) conditioned on the entropy
    h, _ = ctx.mica_sites()
    ctx.null_load(p["mi"], p["hmin"], 3, float(h.max()))
    q = ctx.mica_pairs("hmin", use_null=True)
    g = ctx.null_get()
    assert g["bin_offsets"][-1] == np.sum(p["hmin"] < h.max()) and np.nanmin(q["pvalue"]) > 0


@pytest.mark.parametrize("case", ["dna", "protein"])
def test_permutation_test_vs_oracle(ctx, case):
    """null.method = permutations (miTest, Mica.cpp:92-118): Perm.p.value / Perm.nb of every pair against the oracle's
    own replay of the same shuffle stream.  The early stop hangs on `rep >= mi` between logarithm sums, so a pair whose
    shuffles came within rounding of its MI may stop elsewhere on the host: those are counted and must be such ties."""
    if case == "dna":
        c = H.random_dna_case(26, 60, 11, mean_brlen=0.2, ambiguity=0.05)
        c["codes"][:, 7] = 2; c["codes"][3, 7] = 4            # constant once the unknown character is ignored
        c["codes"][:, 19] = 4                                 # nothing but unknown characters
        max_perm = 300
    else:
        c = H.myoglobin_inputs()
        max_perm = 60
    A = len(c["pi"])
    _setup(ctx, c)
    pv, nb = ctx.mica_permutations(12345, max_perm)
    opv, onb, closest = O.mica_permutations(c["codes"], A, c["code_mask"], 12345, max_perm)
    S = c["codes"].shape[1]
    ii, jj = np.triu_indices(S, 1)
    assert len(pv) == len(ii) and nb.min() >= 0 and nb.max() <= max_perm and pv.min() > 0 and pv.max() <= 1
    diff = (nb != onb) | (pv != opv)
    mi = ctx.mica_pairs("hmin")["mi"]
    assert np.all(closest[diff] <= 1e-12 * np.maximum(1.0, np.abs(mi[diff]))), "a count differs without a rounding tie"
    assert diff.mean() < 0.02, "pairs that differ: %d of %d" % (diff.sum(), len(diff))
    print("permutation test, %s: %d of %d pairs stop elsewhere than the oracle (rounding ties)" % (case, diff.sum(), len(diff)))
    # the stopping rule: 5 shuffles reached the MI, or the budget ran out
    stopped = nb < max_perm
    live = nb > 0
    assert np.all(pv[stopped & live] == 6.0 / (nb[stopped & live] + 1.0))
    if case == "dna":
        const = (ii == 7) | (jj == 7) | (ii == 19) | (jj == 19)
        assert np.all(nb[const] == 0) and np.all(pv[const] == 1.0)
    # same seed, same table; another seed, another stream
    pv2, nb2 = ctx.mica_permutations(12345, max_perm)
    assert np.array_equal(pv, pv2) and np.array_equal(nb, nb2)
    pv3, nb3 = ctx.mica_permutations(777, max_perm)
    assert not np.array_equal(nb, nb3)
    # significant pairs use their whole budget: a column against its copy can only be matched by chance
    if case == "dna":
        c2 = dict(c); c2["codes"] = c["codes"].copy(); c2["codes"][:, 1] = c2["codes"][:, 0]
        _setup(ctx, c2)
        pv4, nb4 = ctx.mica_permutations(5, 200)
        h = ctx.mica_sites()[0]
        if h[0] > 0.8:
            assert nb4[0] == 200 and pv4[0] <= 5.0 / 201
