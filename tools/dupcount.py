import os, sys, hashlib
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
from comap_b200 import api, synthetic as syn
T = 500
parent, brlen = syn.random_tree(T, 20251018, 0.02)
Q, pi = syn.hky85(2.5, [0.3, 0.2, 0.2, 0.3]); rates, probs = syn.gamma_rates(0.5, 4)
ctx = api.Context(device=0)
ctx.set_tree(parent, brlen); ctx.set_model(Q, pi, rates, probs)
n = 400000
codes, cls = ctx.simulate(2, 0, n)
cols = np.ascontiguousarray(codes.T)
h = np.array([hash(c.tobytes()) for c in cols], dtype=np.int64)
u, cnt = np.unique(h, return_counts=True)
const = (cols == cols[:, :1]).all(axis=1)
print("sites", n, "unique", len(u), "constant", const.sum(), "dups beyond constant", n - len(u) - (const.sum() - len(np.unique(cols[const][:, 0]))))
for c in range(4):
    m = cls == c
    print("class", c, "sites", m.sum(), "unique", len(np.unique(h[m])))
