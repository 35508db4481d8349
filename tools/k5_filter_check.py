"""Permutation test of the bench's mica workload (and of a low-entropy alignment, where shuffled tables tie with the observed
one all the time) dumped to an .npz: run once with CMB_K5_FILTER=1 and once with 0, the tables must be identical."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from comap_b200 import api, synthetic as syn  # noqa: E402

out = sys.argv[1]
cfg = bench.MICA
res = {}
ctx = api.Context()
for name, brlen_mean, S in (("bench", cfg["mean_brlen"], cfg["sites"]), ("conserved", 0.01, 300)):
    parent, brlen = syn.random_tree(cfg["taxa"], cfg["tree_seed"], brlen_mean)
    Q, pi = syn.hky85(2.5, [0.3, 0.2, 0.2, 0.3])
    rates, probs = syn.gamma_rates(0.5, 4)
    ctx.set_tree(parent, brlen); ctx.set_model(Q, pi, rates, probs)
    codes, _ = ctx.simulate(cfg["aln_seed"], 0, S)
    ctx.set_alignment(codes, syn.identity_code_mask(4))
    pv, nb = ctx.mica_permutations(cfg["perm_seed"], cfg["max_perm"])
    res[name + "_pv"] = pv; res[name + "_nb"] = nb
    print(name, "pairs", len(pv), "mean shuffles %.1f" % nb.mean(), "pairs at the budget", int((nb == cfg["max_perm"]).sum()))
ctx.close()
np.savez(out, **res)
