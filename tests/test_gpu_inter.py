"""GPU parity of the two-data-set ("inter-gene") analysis -- cmb_pairs_inter / cmb_null_inter,
reference CoETools.cpp:732-840 and AnalysisTools.cpp:662-735 -- against the CPU oracle."""
import numpy as np
import pytest
import helpers as H
import oracle_binding as O
from comap_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
STATS = ["correlation", "covariance", "cosinus", "cosubstitution", "compensation", "corrected_correlation"]
SEED2 = 0x9E3779B97F4A7C15


@pytest.fixture(scope="module")
def two():
    """Two data sets on one topology: different branch lengths, model parameters and lengths."""
    from comap_b200 import api
    a, b = api.Context(), api.Context()
    c1 = H.random_dna_case(18, 70, 21, mean_brlen=0.08)
    c2 = H.random_dna_case(18, 90, 21, mean_brlen=0.08)          # same seed -> same topology
    c2["brlen"] = c2["brlen"] * 1.7
    c2["Q"], c2["pi"] = syn.hky85(4.0, [0.2, 0.3, 0.3, 0.2])
    c2["rates"], c2["probs"] = syn.gamma_rates(1.2, 3)
    rng = np.random.default_rng(77)
    c2["codes"] = H.simulate_np(c2["parent"], c2["brlen"], c2["Q"], c2["pi"], c2["rates"], rng, 90)
    ms = []
    for ctx, c in ((a, c1), (b, c2)):
        ctx.set_tree(c["parent"], c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
        ctx.set_alignment(c["codes"], c["code_mask"])
        ms.append(ctx.map())
    yield a, b, c1, c2, ms[0], ms[1]
    a.close(); b.close()


def _eq(x, y):
    x, y = np.asarray(x), np.asarray(y)
    return np.array_equal(np.isnan(x), np.isnan(y)) and np.array_equal(x[~np.isnan(x)], y[~np.isnan(y)])


@pytest.mark.parametrize("stat", STATS)
def test_rectangle_bit_exact_given_vectors(two, stat):
    a, b, c1, c2, m1, m2 = two
    O.mean_vectors(m1["n"], m2["n"])
    g, k = a.pairs_inter(b, stat)
    o = O.pairs_inter(stat, m1, m2)
    assert k == 70 * 90 == len(o["i"])
    assert np.array_equal(g["i"], o["i"]) and np.array_equal(g["j"], o["j"])   # reference row order
    assert _eq(g["stat"], o["stat"])
    assert np.array_equal(g["rcmin"], o["rcmin"]) and np.array_equal(g["prmin"], o["prmin"])
    assert np.array_equal(g["nmin"], o["nmin"])                                  # upstream's norms2[i] pairing
    g2, _ = a.pairs_inter(b, stat, nmin_by_row=False)
    o2 = O.pairs_inter(stat, m1, m2, nmin_by_row=False)
    assert np.array_equal(g2["nmin"], o2["nmin"]) and not np.array_equal(g2["nmin"], g["nmin"])


def test_rectangle_vs_independent_oracle_mapping(two):
    """End to end against the oracle's own mapping of both data sets (1e-9)."""
    a, b, c1, c2, m1, m2 = two
    q1 = O.map_sites(c1["parent"], c1["brlen"], c1["Q"], c1["pi"], c1["rates"], c1["probs"], c1["codes"], c1["code_mask"])
    q2 = O.map_sites(c2["parent"], c2["brlen"], c2["Q"], c2["pi"], c2["rates"], c2["probs"], c2["codes"], c2["code_mask"])
    g, k = a.pairs_inter(b, "correlation")
    o = O.pairs_inter("correlation", q1, q2)
    ok = ~np.isnan(o["stat"])
    assert np.array_equal(np.isnan(g["stat"]), ~ok)
    assert np.allclose(g["stat"][ok], o["stat"][ok], rtol=1e-9, atol=1e-12)


def test_filters_and_second_data_set_thresholds(two):
    a, b, c1, c2, m1, m2 = two
    f = dict(min_rate_class=1, min_rate=0.2, max_rate_class_diff=1, max_rate_diff=1.5, min_stat=0.05)
    g, k = a.pairs_inter(b, "correlation", filters=f, min_rate_class2=1, min_rate2=0.4)
    o = O.pairs_inter("correlation", m1, m2, min_rate_class1=1, min_rate_class2=1, min_rate1=0.2, min_rate2=0.4,
                      max_rate_class_diff=1, max_rate_diff=1.5, min_stat=0.05)
    assert 0 < k == len(o["i"]) < 70 * 90
    assert np.array_equal(g["i"], o["i"]) and np.array_equal(g["j"], o["j"]) and _eq(g["stat"], o["stat"])


def test_independent_comparisons(two):
    a, b, c1, c2, m1, m2 = two
    with pytest.raises(RuntimeError, match="same length"):
        a.pairs_inter(b, "correlation", independent=True)
    # data set 2 against itself shifted: same length
    from comap_b200 import api
    d = api.Context()
    d.set_tree(c2["parent"], c2["brlen"]); d.set_model(c2["Q"], c2["pi"], c2["rates"], c2["probs"])
    d.set_alignment(np.ascontiguousarray(c2["codes"][:, ::-1]), c2["code_mask"])
    m3 = d.map()
    for stat in ("correlation", "compensation", "corrected_correlation"):
        O.mean_vectors(m2["n"], m3["n"])
        g, k = b.pairs_inter(d, stat, independent=True, filters=dict(min_stat=0.01))
        o = O.pairs_inter(stat, m2, m3, independent=True, min_stat=0.01)
        assert k == len(o["i"]) and np.array_equal(g["i"], o["i"]) and np.array_equal(g["i"], g["j"])
        assert _eq(g["stat"], o["stat"]) and np.array_equal(g["nmin"], o["nmin"])
    d.close()


@pytest.mark.parametrize("stat", ["correlation", "corrected_correlation", "cosubstitution"])
def test_null_inter_equals_exported_alignments(two, stat):
    """cmb_null_inter == the statistic of the mappings of the alignments cmb_simulate exports for
    the two simulators' streams (seed, seed ^ const), bit for bit given the device vectors."""
    from comap_b200 import api
    a, b, c1, c2, m1, m2 = two
    rep_cpu, rep_ram, seed = 3, 150, 4242
    O.mean_vectors(m1["n"], m2["n"])
    raw = a.null_inter(b, stat, seed, rep_cpu, rep_ram)
    mask = syn.identity_code_mask(4)
    x, y = api.Context(), api.Context()
    for ctx, c in ((x, c1), (y, c2)):
        ctx.set_tree(c["parent"], c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
    exp = []
    for i in range(rep_cpu):
        s1, _ = x.simulate(seed, i * rep_ram, rep_ram)
        s2, _ = y.simulate(seed ^ SEED2, i * rep_ram, rep_ram)
        x.set_alignment(s1, mask); y.set_alignment(s2, mask)
        n1, n2 = x.map(), y.map()
        for j in range(rep_ram):
            exp.append((O.stat(stat, n1["n"][j], n2["n"][j]), min(n1["rate_class"][j], n2["rate_class"][j]),
                        min(n1["post_rate"][j], n2["post_rate"][j]), min(n1["norm"][j], n2["norm"][j])))
    exp = np.array(exp)
    assert _eq(raw[:, 0], exp[:, 0])
    assert np.array_equal(raw[:, 1:], exp[:, 1:])
    x.close(); y.close()
