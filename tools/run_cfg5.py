"""Config 5 (BASELINE.json configs[4]): synthetic 20,000-site x 200-taxon protein alignment,
JTT92 + Gamma(4, 1.0), clustering analysis (distance matrix + complete linkage + groups) and
a short clustering null.  Prints device times per kernel family.

    python tools/run_cfg5.py [--sites 20000] [--taxa 200] [--null 2]
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comap_b200 import api, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--sites", type=int, default=20000)
ap.add_argument("--taxa", type=int, default=200)
ap.add_argument("--null", type=int, default=2)
a = ap.parse_args()
parent, brlen = syn.random_tree(a.taxa, 2, 0.05)
Q, pi = syn.jtt92()
rates, probs = syn.gamma_rates(1.0, 4)
ctx = api.Context(device=0)
ctx.set_tree(parent, brlen); ctx.set_model(Q, pi, rates, probs)
t0 = time.time(); codes, _ = ctx.simulate(1, 0, a.sites); t_sim = time.time() - t0
ctx.set_alignment(codes, syn.identity_code_mask(20))
ctx.profile_enable(True); ctx.profile_reset()
t0 = time.time(); m = ctx.map(want_vectors=False); t_map = time.time() - t0
t0 = time.time(); ctx.distance_matrix("correlation", want=False); t_dist = time.time() - t0
t0 = time.time(); left, right, height = ctx.cluster("complete"); t_clu = time.time() - t0
t0 = time.time(); g = ctx.groups("correlation", 10); t_grp = time.time() - t0
t0 = time.time(); nul = ctx.cluster_null("correlation", "complete", 7, 0, a.null, 10) if a.null else None; t_null = time.time() - t0
prof = {k: ctx.profile_get(k) for k in ("map_down", "map_up", "simulate", "distance", "cluster")}
S = a.sites
print(json.dumps(dict(sites=S, taxa=a.taxa, groups=len(g["members"]), wall=dict(simulate=t_sim, map=t_map, distance=t_dist,
      cluster=t_clu, groups=t_grp, null=t_null), device_ms={k: v[0] for k, v in prof.items()},
      heights_sorted=bool(np.all(np.diff(height) >= -1e-12)), max_height=float(height.max()),
      null_rows=None if nul is None else len(nul["rep"]),
      pairs_per_s_clustering=(1 + a.null) * S * (S - 1) / 2 / (t_map + t_dist + t_clu + t_grp + t_null))))
