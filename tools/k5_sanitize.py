"""A small pass over mica's kernels for compute-sanitizer (memcheck / racecheck): column statistics, listed pairs and
the permutation test (shared-memory private copies, warp votes) on a DNA case with ambiguity codes and a protein case."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from comap_b200 import api, synthetic as syn  # noqa: E402

ctx = api.Context()
c = H.random_dna_case(13, 21, 4, mean_brlen=0.2, ambiguity=0.05)
ctx.set_tree(c["parent"], c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
ctx.set_alignment(c["codes"], c["code_mask"])
h, a = ctx.mica_sites()
p = ctx.mica_pairs("hmin")
pv, nb = ctx.mica_permutations(3, 70)
mi, hj = ctx.mica_pair_list(np.array([0, 3, 20]), np.array([5, 3, 1]))
print("dna", len(pv), int(nb.sum()), float(p["mi"].sum()))
T = 9
parent, brlen = syn.random_tree(T, 5, 0.3)
Q, pi = syn.jtt92()
rates, probs = syn.gamma_rates(1.0, 2)
ctx.set_tree(parent, brlen); ctx.set_model(Q, pi, rates, probs)
codes, _ = ctx.simulate(1, 0, 12)
ctx.set_alignment(codes, syn.identity_code_mask(20))
pv, nb = ctx.mica_permutations(4, 40)
p = ctx.mica_pairs("hmin")
print("protein", len(pv), int(nb.sum()), float(p["mi"].sum()))
ctx.close()
