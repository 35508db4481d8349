"""GPU parity of the candidate-groups analysis (cmb_candidates; reference CoMap.cpp:592-711,
CoETools.cpp:901-1087) against a pure-Python restatement of the sampler fed with the same
simulated mappings: group statistics, counters and p-values must be identical."""
import numpy as np
import pytest
import helpers as H
import oracle_binding as O
from comap_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _batches(c, seed, rep_ram):
    """The simulated batches cmb_candidates draws: sites [k R, (k+1) R) of the stream `seed`."""
    from comap_b200 import api
    x = api.Context()
    x.set_tree(c["parent"], c["brlen"]); x.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
    mask = syn.identity_code_mask(4)
    k = 0
    try:
        while True:
            s, _ = x.simulate(seed, k * rep_ram, rep_ram)
            x.set_alignment(s, mask)
            m = x.map()
            yield m["n"], m["norm"]
            k += 1
    finally:
        x.close()


@pytest.mark.parametrize("stat", ["correlation", "compensation", "cosubstitution", "corrected_correlation"])
def test_candidates_equal_reference_sampler(stat):
    from comap_b200 import api
    c = H.random_dna_case(20, 120, 31, mean_brlen=0.1)
    ctx = api.Context()
    ctx.set_tree(c["parent"], c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
    ctx.set_alignment(c["codes"], c["code_mask"])
    m = ctx.map()
    O.mean_vector(m["n"])
    rng = np.random.default_rng(3)
    groups = [list(rng.choice(120, size=k, replace=False)) for k in (2, 3, 2, 4, 2)]
    analysable = [1, 1, 0, 1, 1]                       # one group refers to a site that was filtered out
    kw = dict(omega=0.6, min_sim=40, max_trials=3, rep_ram=200)
    g = ctx.candidates(stat, groups, seed=99, analysable=analysable, **kw)
    o = O.candidates_reference(stat, m["n"], m["norm"], groups, kw["omega"], kw["min_sim"], kw["max_trials"],
                               _batches(c, 99, kw["rep_ram"]), analysable=analysable)
    an = np.array(analysable, bool)
    assert np.array_equal(g["stat"][an], o["stat"][an]) and np.all(np.isnan(g["stat"][~an]))
    assert np.array_equal(g["n2"], o["n2"]) and np.array_equal(g["n1"], o["n1"])
    assert np.array_equal(g["pvalue"][an], o["pvalue"][an]) and np.all(np.isnan(g["pvalue"][~an]))
    assert g["n_simulated"] == o["n_simulated"] and g["n2"][an].max() >= 1
    ctx.close()


def test_candidates_stop_after_max_trials():
    """Norm ranges no simulated site can satisfy: every batch is a failed trial, the loop stops after
    max_trials batches and p = (0+1)/(0+1) (CoETools.cpp:984-989, CoETools.h:226-229)."""
    from comap_b200 import api
    c = H.random_dna_case(10, 50, 5, mean_brlen=0.05)
    ctx = api.Context()
    ctx.set_tree(c["parent"], c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
    ctx.set_alignment(c["codes"], c["code_mask"]); ctx.map()
    g = ctx.candidates("correlation", [[0, 1], [2, 3]], omega=-1.0, min_sim=5, max_trials=4, rep_ram=64, seed=1)
    assert g["n_simulated"] == 4 * 64 and np.all(g["n2"] == 0) and np.all(g["pvalue"] == 1.0)
    with pytest.raises(RuntimeError, match="has 1 sites"):
        ctx.candidates("correlation", [[0]], seed=1)
    ctx.close()
