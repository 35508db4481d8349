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
