"""GPU parity of K2 (pair statistics, p-values), K3 (simulation) and the null pipeline
against the CPU oracle, through the C ABI."""
import numpy as np
import pytest
import helpers as H
import oracle_binding as O
from comap_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
RTOL = 1e-9  # north_star tolerance for pair statistics (fp64)


@pytest.fixture(scope="module")
def ctx():
    from comap_b200 import api
    c = api.Context()
    yield c
    c.close()


def _case(T=24, S=150, seed=11, C=4):
    c = H.random_dna_case(T, S, seed, mean_brlen=0.08, C=C)
    return c


def _setup(ctx, c):
    ctx.set_tree(c["parent"], c["brlen"])
    ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
    ctx.set_alignment(c["codes"], c["code_mask"])
    return ctx.map()


def _close(a, b, rtol=RTOL, atol=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    nan = np.isnan(a)
    return np.array_equal(nan, np.isnan(b)) and np.allclose(a[~nan], b[~nan], rtol=rtol, atol=atol)


def test_simulate_bit_exact_vs_oracle(ctx):
    c = _case(T=40, S=10)
    _setup(ctx, c)
    for weighted in (False, True):
        a, ca = ctx.simulate(1234, 0, 3000, weighted_classes=weighted)
        b, cb = O.simulate(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], 1234, 0, 3000, weighted)
        assert np.array_equal(ca, cb)
        assert np.array_equal(a, b)
    # counter based: a window of the stream equals the same sites simulated alone
    w, _ = ctx.simulate(1234, 1000, 257, weighted_classes=True)
    assert np.array_equal(w, a[:, 1000:1257])


def test_simulate_protein_bit_exact(ctx):
    m = H.myoglobin_inputs()
    ctx.set_tree(m["parent"], m["brlen"]); ctx.set_model(m["Q"], m["pi"], m["rates"], m["probs"])
    a, ca = ctx.simulate(7, 5, 300)
    b, cb = O.simulate(m["parent"], m["brlen"], m["Q"], m["pi"], m["rates"], m["probs"], 7, 5, 300)
    assert np.array_equal(a, b) and np.array_equal(ca, cb)


ALL_STATS = ["correlation", "covariance", "cosinus", "cosubstitution", "compensation", "corrected_correlation"]


@pytest.mark.parametrize("stat", ALL_STATS)
def test_pairs_all_statistics_vs_oracle(ctx, stat):
    c = _case()
    r = _setup(ctx, c)
    q = O.map_sites(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], c["codes"], c["code_mask"])
    O.mean_vector(q["n"])  # the corrected correlation's mean vector (CoMap.cpp:350-359)
    g, k = ctx.pairs(stat, use_null=False)
    o = O.pairs(stat, q["n"], q["norm"], q["post_rate"], q["rate_class"])
    assert k == len(o["i"]) == 150 * 149 // 2
    assert np.array_equal(g["i"], o["i"]) and np.array_equal(g["j"], o["j"])  # reference row order
    assert _close(g["stat"], o["stat"])
    assert np.array_equal(g["rcmin"], o["rcmin"])
    assert np.allclose(g["prmin"], o["prmin"], rtol=RTOL) and np.allclose(g["nmin"], o["nmin"], rtol=RTOL)
    if stat == "cosubstitution":
        assert np.array_equal(g["stat"], o["stat"])  # integer valued: exact


def test_mutual_information_statistic(ctx):
    """statistic=MI(threshold=t) (discretised mutual information, CoETools.cpp:590-595): observed pairs,
    null and p-values against the oracle.  Device and host `log` may differ in the last bit, so the
    statistic is compared at 1e-9; counts of the null per bin are exact (they depend on Nmin only)."""
    c = _case(T=20, S=90, seed=4)
    for thr in (0.99, 0.3):
        ctx.set_mi_threshold(thr); O.set_mi_threshold(thr)
        r = _setup(ctx, c)
        g, k = ctx.pairs("mi", use_null=False)
        o = O.pairs("mi", r["n"], r["norm"], r["post_rate"], r["rate_class"])
        assert k == len(o["i"]) and np.array_equal(g["i"], o["i"]) and np.array_equal(g["j"], o["j"])
        assert np.allclose(g["stat"], o["stat"], rtol=1e-9, atol=1e-15) and (g["stat"] > 1e-3).any()
        rep_cpu, rep_ram, K = 2, 150, 4
        s1 = np.stack([ctx.simulate(5, (2 * i) * rep_ram, rep_ram)[0] for i in range(rep_cpu)])
        s2 = np.stack([ctx.simulate(5, (2 * i + 1) * rep_ram, rep_ram)[0] for i in range(rep_cpu)])
        nmax = float(r["norm"].max())
        raw = ctx.null_intra_from_alignments("mi", s1, s2, K=K, nmax=nmax)
        on = O.null_intra(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], "mi", s1, s2, K, nmax)
        assert np.allclose(raw[:, 0], on["raw"][:, 0], rtol=1e-9, atol=1e-15)
        gn = ctx.null_get()
        assert np.array_equal(gn["bin_offsets"], on["bin_offsets"])
        gp, k2 = ctx.pairs("mi", use_null=True)
        assert k2 == k and np.all((gp["pvalue"][~np.isnan(gp["pvalue"])] > 0) & (gp["pvalue"][~np.isnan(gp["pvalue"])] <= 1))
    ctx.set_mi_threshold(0.99); O.set_mi_threshold(0.99)


def test_pairs_filters_and_shards(ctx):
    c = _case(S=203)
    _setup(ctx, c)
    q = O.map_sites(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], c["codes"], c["code_mask"])
    f = dict(min_rate_class=1, min_rate=0.3, max_rate_class_diff=1, max_rate_diff=1.0, min_stat=0.05)
    g, k = ctx.pairs("correlation", use_null=False, filters=f)
    o = O.pairs("correlation", q["n"], q["norm"], q["post_rate"], q["rate_class"], **f)
    assert k == len(o["i"]) and 0 < k < 203 * 202 // 2
    assert np.array_equal(g["i"], o["i"]) and np.array_equal(g["j"], o["j"])
    assert _close(g["stat"], o["stat"])
    # row shards partition the full result
    full, _ = ctx.pairs("correlation", use_null=False)
    parts = [ctx.pairs("correlation", use_null=False, shard_index=s, shard_count=3)[0] for s in range(3)]
    key = np.concatenate([p["i"].astype(np.int64) * 203 + p["j"] for p in parts])
    order = np.argsort(key, kind="stable")
    assert np.array_equal(key[order], full["i"].astype(np.int64) * 203 + full["j"])
    assert np.array_equal(np.concatenate([p["stat"] for p in parts])[order], full["stat"])
    sizes = [len(p["i"]) for p in parts]
    assert max(sizes) - min(sizes) < 203 * 4  # balanced


def _eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


def test_null_from_alignments_vs_oracle(ctx):
    """Same simulated alignments in -> same null distribution out (1e-9 on the statistics,
    identical bin occupancy)."""
    c = _case(T=16, S=120, seed=5)
    r = _setup(ctx, c)
    q = O.map_sites(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], c["codes"], c["code_mask"])
    rep_cpu, rep_ram, K = 4, 300, 6
    s1 = np.stack([ctx.simulate(99, (2 * i) * rep_ram, rep_ram)[0] for i in range(rep_cpu)])
    s2 = np.stack([ctx.simulate(99, (2 * i + 1) * rep_ram, rep_ram)[0] for i in range(rep_cpu)])
    nmax = float(q["norm"].max())
    assert abs(nmax - r["norm"].max()) <= 1e-12 * nmax
    raw = ctx.null_intra_from_alignments("correlation", s1, s2, K=K, nmax=nmax)
    o = O.null_intra(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], "correlation", s1, s2, K, nmax)
    assert _close(raw[:, 0], o["raw"][:, 0])
    assert np.array_equal(raw[:, 1], o["raw"][:, 1])
    assert np.allclose(raw[:, 2:], o["raw"][:, 2:], rtol=RTOL)
    g = ctx.null_get()
    assert g["K"] == K and np.array_equal(g["bin_offsets"], o["bin_offsets"])
    fin = ~np.isnan(o["sorted"])
    assert _close(g["sorted"][fin], o["sorted"][fin])


@pytest.mark.parametrize("stat", ALL_STATS)
def test_statistics_and_pvalues_bit_exact_given_vectors(ctx, stat):
    """north_star: null counts and p-values bit-exact.  p = (nsim - #{sim < stat} + 1)/(nsim + 1)
    with massive ties among low-rate sites (r = 1 between constant sites), so one ulp moves
    counts by dozens.  The pair kernels therefore replicate the reference's operation order
    without fused multiply-adds: fed the SAME mapping vectors, statistics, bin occupancy,
    counts and p-values equal the CPU restatement bit for bit."""
    c = _case(T=16, S=120, seed=5)
    r = _setup(ctx, c)
    rep_cpu, rep_ram, K = 3, 200, 6
    s1 = np.stack([ctx.simulate(99, (2 * i) * rep_ram, rep_ram)[0] for i in range(rep_cpu)])
    s2 = np.stack([ctx.simulate(99, (2 * i + 1) * rep_ram, rep_ram)[0] for i in range(rep_cpu)])
    nmax = 0.6 * float(r["norm"].max())  # some pairs fall outside [0, nmax) -> "NA\t0" rows
    O.mean_vector(r["n"])  # observed mean vector, used for observed AND simulated pairs
    raw = ctx.null_intra_from_alignments(stat, s1, s2, K=K, nmax=nmax)
    g = ctx.null_get()
    gp, k = ctx.pairs(stat, use_null=True)
    # device mapping vectors of the simulated batches, then the CPU restatement on them
    mask = syn.identity_code_mask(4)
    exp_stat, exp_nmin = [], []
    for i in range(rep_cpu):
        ctx.set_alignment(s1[i], mask); m1 = ctx.map()
        ctx.set_alignment(s2[i], mask); m2 = ctx.map()
        for j in range(rep_ram):
            exp_stat.append(O.stat(stat, m1["n"][j], m2["n"][j]))
            exp_nmin.append(min(np.sqrt((m1["n"][j] ** 2).sum()), np.sqrt((m2["n"][j] ** 2).sum())))
    assert _eq(raw[:, 0], exp_stat)
    cat = np.array([O.domain_index(0, nmax, K, x) for x in raw[:, 3]])
    assert np.array_equal(np.diff(g["bin_offsets"]), np.bincount(cat[cat >= 0], minlength=K))
    op = O.pairs(stat, r["n"], r["norm"], r["post_rate"], r["rate_class"], null=(K, nmax, g["bin_offsets"], g["sorted"]))
    assert k == len(op["i"])
    assert _eq(gp["stat"], op["stat"])
    assert np.array_equal(gp["nsim"], op["nsim"])
    assert _eq(gp["pvalue"], op["pvalue"])
    assert np.array_equal(gp["nmin"], op["nmin"]) and np.array_equal(gp["prmin"], op["prmin"])
    assert (gp["nsim"] == 0).any() and np.all(np.isnan(gp["pvalue"][gp["nsim"] == 0]))  # "NA\t0" rows


def test_null_intra_device_rng_equals_exported_alignments(ctx):
    """cmb_null_intra (device RNG) == cmb_null_intra_from_alignments on the alignments that
    cmb_simulate exports for the same global site indices; shards concatenate."""
    c = _case(T=16, S=60, seed=8)
    _setup(ctx, c)
    rep_cpu, rep_ram, K = 5, 128, 4
    raw = ctx.null_intra("correlation", 4242, rep_cpu, rep_ram, K=K, nmax=2.0, want_raw=True)
    s1 = np.stack([ctx.simulate(4242, (2 * i) * rep_ram, rep_ram)[0] for i in range(rep_cpu)])
    s2 = np.stack([ctx.simulate(4242, (2 * i + 1) * rep_ram, rep_ram)[0] for i in range(rep_cpu)])
    raw2 = ctx.null_intra_from_alignments("correlation", s1, s2, K=K, nmax=2.0)
    assert np.array_equal(np.isnan(raw), np.isnan(raw2))
    assert np.array_equal(np.nan_to_num(raw), np.nan_to_num(raw2))
    a = ctx.null_intra("correlation", 4242, rep_cpu, rep_ram, K=0, rep_begin=0, rep_end=2, want_raw=True)
    b = ctx.null_intra("correlation", 4242, rep_cpu, rep_ram, K=0, rep_begin=2, rep_end=5, want_raw=True)
    assert np.array_equal(np.nan_to_num(np.concatenate([a, b])), np.nan_to_num(raw))


def test_null_other_statistics(ctx):
    c = _case(T=12, S=40, seed=9)
    _setup(ctx, c)
    rep_cpu, rep_ram = 2, 200
    s1 = np.stack([ctx.simulate(1, (2 * i) * rep_ram, rep_ram)[0] for i in range(rep_cpu)])
    s2 = np.stack([ctx.simulate(1, (2 * i + 1) * rep_ram, rep_ram)[0] for i in range(rep_cpu)])
    for stat in ("cosubstitution", "compensation", "cosinus", "covariance"):
        raw = ctx.null_intra_from_alignments(stat, s1, s2, K=3, nmax=3.0)
        o = O.null_intra(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], stat, s1, s2, 3, 3.0)
        assert _close(raw[:, 0], o["raw"][:, 0]), stat
        assert np.array_equal(ctx.null_get()["bin_offsets"], o["bin_offsets"])


def test_pair_columns_in_two_calls_overlapping_the_null(ctx):
    """Null-independent columns scored and fetched first, PValue / Nsim after the null: same
    table as one call (the end-to-end flow of bench.py)."""
    c = _case(S=177)
    _setup(ctx, c)
    n0 = ctx.pairs_resident("correlation", use_null=False, columns=0x3F)
    bufs = [np.empty(n0, dt) for dt in ctx.COL_DTYPE]
    for k in range(6):
        ctx.pairs_fetch(k, bufs[k])
    ctx.null_intra("correlation", 5, 3, 160, K=5)
    n1 = ctx.pairs_resident("correlation", use_null=True, columns=0xC0)
    for k in (6, 7):
        ctx.pairs_fetch(k, bufs[k])
    ctx.sync()
    ref, k = ctx.pairs("correlation", use_null=True)
    assert n0 == n1 == k == 177 * 176 // 2
    for name, b in zip(ctx.COLS, bufs):
        assert _eq(b, ref[name]), name
    with pytest.raises(RuntimeError):          # a column that was never produced cannot be fetched
        ctx.map()
        ctx.pairs_resident("correlation", use_null=False, columns=0x04)
        ctx.pairs_fetch(0, bufs[0])


def test_deferred_mapping_overlaps_and_equals_the_blocking_one(ctx):
    """cmb_map with every output NULL is only enqueued (side stream); the null and the pair table computed
    after it equal the blocking flow bit for bit, and a saturated alignment is reported by the first call
    that needs the mapping."""
    c = _case(S=211)
    ctx.set_tree(c["parent"], c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
    ctx.set_alignment(c["codes"], c["code_mask"])
    ctx.map()
    ctx.null_intra("correlation", 5, 3, 160, K=5)
    ref, k = ctx.pairs("correlation", use_null=True)
    ctx.set_alignment(c["codes"], c["code_mask"])
    ctx.map_async()
    ctx.null_intra("correlation", 5, 3, 160, K=5)      # nmax = -1: completes the mapping for its max norm
    got, k2 = ctx.pairs("correlation", use_null=True)
    assert k == k2
    for name in ctx.COLS:
        assert _eq(got[name], ref[name]), name
    ctx.map_async(); ctx.map_async()                    # re-enqueueing waits for the one in flight
    with pytest.raises(RuntimeError):                   # the null was binned with the previous mapping's max norm
        ctx.pairs("correlation", use_null=True)
    got, _ = ctx.pairs("correlation", use_null=False)
    assert _eq(got["stat"], ref["stat"])


@pytest.mark.parametrize("kind,alpha,pinv", [("gamma", 0.5, 0.0), ("gamma", 2.0, 0.0), ("invariant", 1.0, 0.3), ("constant", 1.0, 0.0)])
def test_continuous_rate_simulation_vs_oracle(ctx, kind, alpha, pinv):
    """simulations.continuous = yes (CoMap.cpp:146,209-219): one continuous rate per site, P(d r) on the fly.
    Device and oracle run the same sampler on the same Philox stream; their exp / log / cos differ in the last
    bits, so a draw that lands within an ulp of a cumulative probability may differ: at most 1 column in 1000."""
    c = _case(T=40, S=10)
    _setup(ctx, c)
    try:
        ctx.set_continuous_rates(kind, alpha, pinv)
        a, cls = ctx.simulate(321, 7, 4000)
        b, rates = O.simulate_continuous(c["parent"], c["brlen"], c["Q"], c["pi"], kind, alpha, pinv, 321, 7, 4000)
        assert np.all(cls == -1)
        differing = (a != b).any(axis=0).mean()
        assert differing <= 1e-3, differing
        w, _ = ctx.simulate(321, 1007, 100)          # counter based: a window equals the same sites alone
        assert np.array_equal(w, a[:, 1000:1100])
        # the null distribution runs on the continuous simulator and equals the exported alignments' null
        raw = ctx.null_intra("correlation", 5, 2, 128, K=3, nmax=2.0, want_raw=True)
        s1 = np.stack([ctx.simulate(5, (2 * i) * 128, 128)[0] for i in range(2)])
        s2 = np.stack([ctx.simulate(5, (2 * i + 1) * 128, 128)[0] for i in range(2)])
        raw2 = ctx.null_intra_from_alignments("correlation", s1, s2, K=3, nmax=2.0)
        assert np.array_equal(np.nan_to_num(raw), np.nan_to_num(raw2))
    finally:
        ctx.set_continuous_rates("off")
    d, cls = ctx.simulate(321, 7, 50)                # back to the discrete classes
    e, _ = O.simulate(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], 321, 7, 50)
    assert np.array_equal(d, e) and np.all(cls >= 0)


def test_continuous_rate_simulation_protein(ctx):
    m = H.myoglobin_inputs()
    ctx.set_tree(m["parent"], m["brlen"]); ctx.set_model(m["Q"], m["pi"], m["rates"], m["probs"])
    try:
        ctx.set_continuous_rates("gamma", 0.985435)
        a, _ = ctx.simulate(7, 5, 600)
        b, r = O.simulate_continuous(m["parent"], m["brlen"], m["Q"], m["pi"], "gamma", 0.985435, 0.0, 7, 5, 600)
        assert (a != b).any(axis=0).mean() <= 5e-3
    finally:
        ctx.set_continuous_rates("off")


def test_pattern_compression_of_the_null_is_bit_identical(ctx, monkeypatch):
    """Constant simulated columns are mapped once per state and batch (as Bio++ maps distinct site patterns):
    raw null rows, bin occupancy and the sorted null equal the uncompressed run bit for bit -- on a slow model
    where most columns are constant, and across several batches."""
    c = H.random_dna_case(30, 64, 19, mean_brlen=0.01, C=4)
    _setup(ctx, c)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("CMB_NULL_DEDUP", flag)
        ctx.profile_reset()
        raw = ctx.null_intra("correlation", 77, 6, 700, K=4, nmax=1.0, want_raw=True)
        g = ctx.null_get()
        out[flag] = (raw, g, ctx.profile_get("sites_simulated")[0], ctx.profile_get("sites_mapped_null")[0])
    a, b = out["0"], out["1"]
    assert np.array_equal(np.isnan(a[0]), np.isnan(b[0])) and np.array_equal(np.nan_to_num(a[0]), np.nan_to_num(b[0]))
    assert np.array_equal(a[1]["bin_offsets"], b[1]["bin_offsets"])
    assert np.array_equal(np.nan_to_num(a[1]["sorted"]), np.nan_to_num(b[1]["sorted"]))
    assert a[2] == b[2] == 2 * 6 * 700 and a[3] == a[2] and b[3] < 0.8 * b[2]    # most columns were constant


def test_label_mutual_information_pairs_null_and_pvalues(ctx):
    """statistic=MI with nijt=Label, nijt.average=no (CoETools.cpp:577-589): labels mapped on the device, the
    statistic of every pair against the oracle given the same vectors (1e-9: device and host log differ in the last
    bit), the null built from exported alignments against the oracle running the same variant, p-values
    self-consistent against the device's own sorted null."""
    c = _case(T=18, S=150, seed=21)
    try:
        ctx.set_tree(c["parent"], c["brlen"])
        ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"], count_method="label")
        ctx.set_alignment(c["codes"], c["code_mask"])
        ctx.set_map_mode(False, True)
        r = ctx.map()
        O.set_map_mode(False, True); O.set_mi_label(4)
        g, k = ctx.pairs("mi_label", use_null=False)
        o = O.pairs("mi_label", r["n"], r["norm"], r["post_rate"], r["rate_class"])
        assert k == len(o["i"]) == 150 * 149 // 2
        assert np.allclose(g["stat"], o["stat"], rtol=1e-9, atol=1e-12) and g["stat"].max() > 0.05
        # sites with more substitutions than the kernel's in-register list take the column re-reading path
        dense = np.tile(np.arange(1, 13, dtype=float), 20)[: r["n"].shape[1]]
        n2 = r["n"].copy(); n2[0] = dense; n2[1] = dense[::-1]; n2[2] = np.roll(dense, 5)
        ctx.load_vectors(n2)
        g2, _ = ctx.pairs("mi_label", use_null=False)
        nrm = np.sqrt((n2 ** 2).sum(1))
        o2 = O.pairs("mi_label", n2, nrm, r["post_rate"], r["rate_class"])
        assert np.allclose(g2["stat"], o2["stat"], rtol=1e-9, atol=1e-12)
        ctx.map()
        # null: same exported alignments -> same statistics and bins
        rep_cpu, rep_ram, K = 3, 200, 4
        s1 = np.stack([ctx.simulate(5, (2 * i) * rep_ram, rep_ram)[0] for i in range(rep_cpu)])
        s2 = np.stack([ctx.simulate(5, (2 * i + 1) * rep_ram, rep_ram)[0] for i in range(rep_cpu)])
        nmax = float(r["norm"].max())
        raw = ctx.null_intra_from_alignments("mi_label", s1, s2, K=K, nmax=nmax)
        on = O.null_intra(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], "mi_label", s1, s2, K, nmax,
                          method="label")
        same = np.isclose(raw[:, 0], on["raw"][:, 0], rtol=1e-9, atol=1e-12)
        assert same.mean() > 0.995                       # an arg-max tie may pick another label at a site
        assert np.allclose(raw[same, 3], on["raw"][same, 3], rtol=1e-9)
        # device RNG path == exported alignments, and p-values against the device's own null
        raw2 = ctx.null_intra("mi_label", 5, rep_cpu, rep_ram, K=K, nmax=nmax, want_raw=True)
        assert np.array_equal(raw2, raw)
        gn = ctx.null_get()
        gp, k2 = ctx.pairs("mi_label", use_null=True)
        assert k2 == k and np.array_equal(gp["stat"], g["stat"])
        # p = (nsim - #{sim < stat} + 1) / (nsim + 1) in the bin of Nmin (CoETools.cpp:700-716), from the device's own
        # statistics and sorted null: a discrete statistic ties EXACTLY with null samples, and device and host log
        # differ in the last bit, so the oracle's statistics would break those ties differently
        off, srt = gn["bin_offsets"], gn["sorted"]
        for t in range(k):
            cat = O.domain_index(0.0, gn["nmax"], K, gp["nmin"][t])
            if cat < 0:
                assert np.isnan(gp["pvalue"][t]) and gp["nsim"][t] == 0
                continue
            b = srt[off[cat]:off[cat + 1]]
            cnt = int(np.searchsorted(b, gp["stat"][t], side="left"))
            assert gp["nsim"][t] == len(b) and gp["pvalue"][t] == (len(b) - cnt + 1) / (len(b) + 1)
    finally:
        O.set_map_mode(); ctx.set_map_mode()
