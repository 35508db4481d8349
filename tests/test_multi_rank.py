"""N > 1 host logic on CPU: world_size-2 (and 3) gloo runs of the sharding + exchange step
(comap_b200/parallel.py).  The compute of every shard is done by the CPU oracle here (the
product path needs a GPU); what is under test is that shards + one all-gather reproduce the
single-process null distribution exactly and that the pair rows are dealt completely,
disjointly and evenly."""
import os
import socket
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from comap_b200 import parallel as par, synthetic as syn  # noqa: E402


def test_replicate_bounds_and_rows():
    for rep, world in ((1000, 8), (7, 3), (2, 4), (0, 2)):
        b = par.replicate_bounds(rep, world)
        assert b[0][0] == 0 and b[-1][1] == rep and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        sizes = [e - s for s, e in b]
        assert max(sizes) - min(sizes) <= 1
    for S, world in ((5000, 8), (129, 2), (10, 4), (7, 3)):
        rows = [par.owned_rows(S, r, world) for r in range(world)]
        allr = np.sort(np.concatenate(rows))
        assert np.array_equal(allr, np.arange(S))                          # complete and disjoint
        pairs = [par.owned_pairs(S, r, world) for r in range(world)]
        assert sum(pairs) == S * (S - 1) // 2
        if S >= 100 * world:
            assert max(pairs) - min(pairs) <= 2 * world * S // 10 and max(pairs) / min(pairs) < 1.02   # balanced


def _case():
    parent, brlen = syn.random_tree(8, 5, 0.1)
    Q, pi = syn.hky85(2.0, [0.3, 0.2, 0.2, 0.3])
    rates, probs = syn.gamma_rates(0.6, 3)
    return parent, brlen, Q, pi, rates, probs


def _shard_null(r0, r1, rep_ram, seed):
    import oracle_binding as O
    parent, brlen, Q, pi, rates, probs = _case()
    if r1 == r0:
        return np.zeros(0), np.zeros(0)
    s1 = np.stack([O.simulate(parent, brlen, Q, pi, rates, probs, seed, (2 * i) * rep_ram, rep_ram)[0] for i in range(r0, r1)])
    s2 = np.stack([O.simulate(parent, brlen, Q, pi, rates, probs, seed, (2 * i + 1) * rep_ram, rep_ram)[0] for i in range(r0, r1)])
    o = O.null_intra(parent, brlen, Q, pi, rates, probs, "correlation", s1, s2, 4, 1.0)
    return o["raw"][:, 0].copy(), o["raw"][:, 3].copy()


def _worker(rank, world, port, rep_cpu, rep_ram, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bounds = par.replicate_bounds(rep_cpu, world)
        r0, r1 = bounds[rank]
        stat, nmin = _shard_null(r0, r1, rep_ram, 9)
        max_per = max(e - s for s, e in bounds) * rep_ram
        st, nm = par.all_gather_null(torch.from_numpy(stat), torch.from_numpy(nmin), len(stat), max_per)
        np.save(os.path.join(out_dir, "stat%d.npy" % rank), st.numpy())
        np.save(os.path.join(out_dir, "nmin%d.npy" % rank), nm.numpy())
        t = par.max_over_ranks(float(rank + 1), "cpu")
        assert t == float(world)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,rep_cpu", [(2, 5), (3, 4), (2, 1)])
def test_sharded_null_equals_single_process(world, rep_cpu, tmp_path):
    import torch.multiprocessing as mp
    rep_ram = 24
    mp.spawn(_worker, args=(world, _free_port(), rep_cpu, rep_ram, str(tmp_path)), nprocs=world, join=True)
    ref_stat, ref_nmin = _shard_null(0, rep_cpu, rep_ram, 9)
    got = [np.load(os.path.join(str(tmp_path), "stat%d.npy" % r)) for r in range(world)]
    gotn = [np.load(os.path.join(str(tmp_path), "nmin%d.npy" % r)) for r in range(world)]
    for r in range(1, world):                                               # every rank holds the same union
        assert np.array_equal(got[0], got[r], equal_nan=True) and np.array_equal(gotn[0], gotn[r], equal_nan=True)
    # padding slots are NaN in both arrays; the rest is the single-process sample list in replicate order
    pad = np.isnan(gotn[0])
    assert pad.sum() == len(got[0]) - rep_cpu * rep_ram
    assert np.array_equal(got[0][~pad], ref_stat, equal_nan=True) and np.array_equal(gotn[0][~pad], ref_nmin)
    # binning the union (NaN dropped like out-of-domain samples) gives the single-process bins
    K, nmax = 4, float(ref_nmin.max()) * 1.0000001
    def bins(st, nm):
        ok = ~np.isnan(nm) & ~np.isnan(st) & (nm < nmax)
        cat = np.minimum((nm[ok] / (nmax / K)).astype(int), K - 1)
        return [np.sort(st[ok][cat == k]) for k in range(K)]
    for a, b in zip(bins(got[0], gotn[0]), bins(ref_stat, ref_nmin)):
        assert np.array_equal(a, b)
