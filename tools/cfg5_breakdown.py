"""Wall-clock breakdown of one configs[4] step (host + device) per API call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from comap_b200 import api, synthetic as syn
S, T = 20000, 200
parent, brlen = syn.random_tree(T, 2, 0.05)
Q, pi = syn.jtt92(); rates, probs = syn.gamma_rates(1.0, 4)
ctx = api.Context(device=0)
ctx.set_tree(parent, brlen); ctx.set_model(Q, pi, rates, probs)
codes, _ = ctx.simulate(1, 0, S)
mask = syn.identity_code_mask(20)
def step(verbose):
    t = [time.time()]
    ctx.set_alignment(codes, mask); t.append(time.time())
    ctx.map(want_vectors=False); t.append(time.time())
    ctx.distance_matrix("correlation", want=False); t.append(time.time())
    ctx.cluster("complete"); t.append(time.time())
    ctx.groups("correlation", 10); t.append(time.time())
    ctx.cluster_null("correlation", "complete", 7, 0, 4, 10); t.append(time.time())
    if verbose:
        names = ["set_alignment", "map", "distance", "cluster", "groups", "cluster_null(4)"]
        print(" | ".join("%s %.1f ms" % (n, 1e3 * (b - a)) for n, a, b in zip(names, t[:-1], t[1:])), "| total %.1f ms" % (1e3 * (t[-1] - t[0])))
step(False); step(True); step(True)
