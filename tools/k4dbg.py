import sys, os, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
from comap_b200 import api, synthetic as syn
S, T = 20000, 200
parent, brlen = syn.random_tree(T, 2, 0.05)
Q, pi = syn.jtt92(); rates, probs = syn.gamma_rates(1.0, 4)
ctx = api.Context(device=0)
ctx.set_tree(parent, brlen); ctx.set_model(Q, pi, rates, probs)
codes, _ = ctx.simulate(1, 0, S)
ctx.set_alignment(codes, syn.identity_code_mask(20)); ctx.map(want_vectors=False)
ctx.distance_matrix("correlation", want=False)
t0 = time.time(); l, r, h = ctx.cluster("complete"); print("cluster wall", time.time() - t0, file=sys.stderr)
