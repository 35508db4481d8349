"""CPU checks of the measurement plumbing in bench.py: the order-independent table checksum that lets the scaling
lines carry correctness, the parser of the committed ncu summaries behind `roofline.traffic`, the replicate
ranges the library and the launcher must agree on, and the CPU-baseline model's validation run."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from comap_b200 import parallel as par  # noqa: E402


def _table(n, seed):
    rng = np.random.default_rng(seed)
    i = rng.integers(0, 5000, n).astype(np.int32); j = rng.integers(0, 5000, n).astype(np.int32)
    stat = rng.standard_normal(n); pv = rng.random(n); pv[rng.random(n) < 0.05] = np.nan
    cols = [i, j, stat, np.zeros(n, np.int32), np.zeros(n), np.zeros(n), pv, rng.integers(0, 1000, n).astype(np.int32)]
    return cols


def test_table_checksum_is_shard_additive_and_order_independent():
    cols = _table(10000, 1)
    whole = bench.table_checksum(cols, 10000)
    perm = np.random.default_rng(2).permutation(10000)
    assert bench.table_checksum([c[perm] for c in cols], 10000) == whole
    parts = [bench.table_checksum([c[a:b] for c in cols], b - a) for a, b in ((0, 3000), (3000, 3001), (3001, 10000))]
    assert sum(parts) % (1 << 64) == whole
    changed = [c.copy() for c in cols]
    changed[2][77] = np.nextafter(changed[2][77], 1.0)     # one statistic moved by an ulp
    assert bench.table_checksum(changed, 10000) != whole
    changed = [c.copy() for c in cols]
    changed[7][5] += 1                                     # one Nsim off by one
    assert bench.table_checksum(changed, 10000) != whole


def test_profile_parser_reads_the_committed_summaries():
    d = bench.profile_dram_bytes("r2*_k1_up_mma.txt")
    assert d is not None and d["file"].startswith("profiles/") and d["grid"] > 0 and d["ms"] > 0
    per_site = d["dram_bytes"] / (d["grid"] * d["sites_per_cta"])
    assert 30e3 < per_site < 80e3      # inner partials read once + output rows: tens of KB per site at config 4
    assert bench.profile_dram_bytes("no_such_profile_*.txt") is None
    p = bench.load_peaks()
    assert p["fp64_dmma_tflops"] > 30 and p["hbm_gbs"] > 1000


def test_replicate_ranges_match_the_library_rule():
    """cmb_null_intra_sharded: contiguous ranges, the first rep_cpu % n ranks hold one more (capi_stats.cu
    replicate_range); parallel.replicate_bounds states the same rule for the launcher."""
    for rep, n in ((1000, 8), (7, 3), (2, 4), (0, 2), (13, 5)):
        q, m = divmod(rep, n)
        want = []
        for r in range(n):
            b = r * q + min(r, m)
            want.append((b, b + q + (1 if r < m else 0)))
        assert par.replicate_bounds(rep, n) == want


def test_cpu_model_validation_on_a_small_job():
    cfg = dict(bench.CFG); cfg.update(taxa=30, sites=200)
    w = bench.workload(cfg)
    one = bench.cpu_sample(cfg, w, sample_sites=80, sample_ram=200, seed=3)
    assert set(one["fit"]) == {"per_null_site", "per_obs_site", "per_pair_base", "per_pair_scan_per_sample"}
    v = bench.cpu_validation(cfg, w, one["fit"], sites=100, rep_cpu=3, rep_ram=200)
    assert v["measured_seconds"] > 0 and 0.3 < v["measured_over_predicted"] < 3.0


def test_cpu_sample_keeps_the_composition_of_the_step():
    """The CPU arm scores a bounded sample with the step's own ratio of null pairs to observed pairs, so its value is
    sample pairs / sample seconds (measured), not an extrapolation."""
    cfg = dict(bench.CFG); cfg.update(taxa=20, sites=300, rep_cpu=10, rep_ram=100)
    w = bench.workload(cfg)
    r = bench.cpu_sample(cfg, w, sample_sites=60, seed=1)
    total = 300 * 299 // 2 + 1000
    assert abs(r["sample_fraction"] - (60 * 59 // 2 + round(1000 * (60 * 59 / 2) / (300 * 299 / 2))) / total) < 1e-12  # 1770 + 39 pairs
    assert r["value"] == r["sample_pairs"] / r["sample_seconds"] and r["full_step_seconds"] == total / r["value"]
    assert 0.2 < r["extrapolated_value"] / r["value"] < 5.0


def test_profile_metric_reads_the_committed_k5_summary():
    """bench.py --workload mica takes the permutation kernel's warp-instruction count and DRAM bytes from the newest
    committed ncu summary: the parser finds them, and they describe an issue-bound kernel that moves no memory."""
    inst, src = bench.profile_metric("r2*_k5_permutations.txt", "smsp__inst_executed.sum")
    assert src and src.startswith("profiles/") and inst > 1e9
    dr = bench.profile_dram_bytes("r2*_k5_permutations.txt")
    assert dr and dr["dram_bytes"] < 1e6 and dr["ms"] > 1.0
    assert bench.profile_metric("no_such_kernel_*.txt", "smsp__inst_executed.sum") == (None, None)
