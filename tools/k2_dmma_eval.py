"""Evidence for the K2 choice: the unfused DMUL + DADD Gram tiles (reference's summation order)
against the opt-in DMMA tiles (CMB_K2_DMMA=1) on BASELINE configs[3]: kernel time, and how many statistics /
p-values of the 12 497 500 pairs change when the SAME vectors and the SAME null are scored by the tensor cores.

    python tools/k2_dmma_eval.py > profiles/r2_k2_dmma_eval.json
"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comap_b200 import api, synthetic as syn

S, T, RC, R, K = 5000, 500, 200, 1000, 10
parent, brlen = syn.random_tree(T, 20251018, 0.02)
Q, pi = syn.hky85(2.5, [0.3, 0.2, 0.2, 0.3]); rates, probs = syn.gamma_rates(0.5, 4)
ctx = api.Context(device=0)
ctx.set_tree(parent, brlen); ctx.set_model(Q, pi, rates, probs)
codes, _ = ctx.simulate(1, 0, S)
ctx.set_alignment(codes, syn.identity_code_mask(4)); ctx.map(want_vectors=False)
ctx.null_intra("correlation", 2, RC, R, K=K)
out = {}
tabs = {}
for name, flag in (("unfused", "0"), ("dmma", "1")):
    os.environ["CMB_K2_DMMA"] = flag
    for _ in range(3):
        ctx.pairs_resident("correlation", use_null=True)
    ctx.sync(); ctx.profile_reset(); ctx.profile_enable(True)
    for _ in range(5):
        ctx.pairs_resident("correlation", use_null=True)
    ms, n = ctx.profile_get("pairs")
    ctx.profile_enable(False)
    tabs[name], k = ctx.pairs("correlation", use_null=True)
    flops = 2.0 * (2 * T - 3) * k
    out[name] = dict(ms_per_launch=ms / 5, tflops=flops / (ms / 5 * 1e-3) / 1e12, pairs=k)
a, b = tabs["unfused"], tabs["dmma"]
fin = ~np.isnan(a["stat"])
rel = np.abs(a["stat"][fin] - b["stat"][fin]) / np.maximum(np.abs(a["stat"][fin]), 1e-300)
pv = ~np.isnan(a["pvalue"])
out["difference"] = dict(stat_rows=int(fin.sum()), stat_bit_identical=int((a["stat"][fin] == b["stat"][fin]).sum()),
                         stat_max_rel_diff=float(rel.max()), pvalue_rows=int(pv.sum()),
                         pvalue_mismatches=int((a["pvalue"][pv] != b["pvalue"][pv]).sum()),
                         nsim_identical=bool(np.array_equal(a["nsim"], b["nsim"])),
                         largest_pvalue_shift=float(np.abs(a["pvalue"][pv] - b["pvalue"][pv]).max()))
out["config"] = dict(sites=S, taxa=T, null="%dx%d" % (RC, R), bins=K)
print(json.dumps(out, indent=1))
