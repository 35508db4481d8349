"""Pins the CPU oracle against the reference's only committed numerical outputs
(examples/Proteins/Benchmark/CoMap/Myo_{unif,decomp}.vec and Myo.infos).

The goldens were written in 2012 by CoMap with 6 significant digits.  logLn agrees to the
printed precision; mapping vectors agree to 2e-6 median / <1e-4 max relative (residual =
older Bio++ gamma-quantile/eigen precision; any wrong JTT92 entry would show at 1e-3)."""
import numpy as np
import helpers as H
import oracle_binding as O


def _run(method):
    m = H.myoglobin_inputs()
    r = O.map_sites(m["parent"], m["brlen"], m["Q"], m["pi"], m["rates"], m["probs"], m["codes"],
                    m["code_mask"], method=method)
    return m, r


def test_site_selection_matches_golden_header():
    m = H.myoglobin_inputs()
    assert m["codes"].shape == (100, 129)
    assert np.array_equal(m["coords"], m["golden"]["vec_coords"])
    assert np.array_equal(m["coords"], m["golden"]["infos_coord"])
    # branch order = post-order node ids; golden 'Mean' column = branch lengths
    assert np.allclose(m["brlen"][:-1], m["golden"]["vec_brlen"], rtol=1e-5)


def test_mapping_uniformization_vs_golden():
    m, r = _run("uniformization")
    gold = m["golden"]["vec_unif"].T  # [site][branch]
    rel = np.abs(r["n"] - gold) / np.abs(gold)
    assert np.median(rel) < 5e-6
    assert rel.max() < 1e-4
    assert (rel > 1e-5).mean() < 0.1


def test_mapping_decomposition_vs_golden():
    m, r = _run("decomposition")
    gold = m["golden"]["vec_decomp"].T
    rel = np.abs(r["n"] - gold) / np.abs(gold)
    # on the 1e-6 branches the eigen closed form loses digits on the O(1e-17) diagonal
    # counts (values ~3e-13); bound those absolutely instead
    big = gold > 1e-9
    assert np.median(rel) < 5e-6 and rel[big].max() < 1e-4
    assert np.abs(r["n"] - gold)[~big].max() < 1e-13


def test_uniformization_equals_decomposition():
    m, r1 = _run("uniformization")
    _, r2 = _run("decomposition")
    assert np.allclose(r1["n"], r2["n"], rtol=1e-8, atol=1e-15)


def test_infos_vs_golden():
    m, r = _run("uniformization")
    g = m["golden"]
    # 6 significant digits printed
    assert np.all(np.abs(r["loglik"] - g["infos_logl"]) <= 6e-6 * np.abs(g["infos_logl"]))
    assert np.all(np.abs(r["post_rate"] - g["infos_pr"]) <= 2e-5 * np.abs(g["infos_pr"]))
    assert np.array_equal(r["rate_class"], g["infos_rc"])
    # IsComplete = 0 exactly where an ambiguity code occurs
    amb = (m["codes"] >= 20).any(0)
    assert np.array_equal(~amb, g["infos_complete"].astype(bool))
    assert not g["infos_const"].any()
