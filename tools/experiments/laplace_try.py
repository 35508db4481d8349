# scratch: can a Taylor-series ("Laplace") count reproduce Myo_laplace.vec?
import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import helpers as H
sys.path.insert(0, 'tests/golden')
from make_golden import load_vec
from math import factorial
c = H.myoglobin_inputs()
coords, mean, G = load_vec('/root/reference/examples/Proteins/Benchmark/CoMap/Myo_laplace.vec')
U = c["golden"]["vec_unif"]
print("golden laplace vs unif max abs", np.abs(G - U).max())

def make(variant, trunc):
    def counts(Q, pi, T, weights=None):
        A = len(pi)
        QL = Q.copy(); np.fill_diagonal(QL, 0)
        pw = [np.linalg.matrix_power(Q, p) for p in range(trunc + 1)]
        m = np.zeros((A, A))
        hi = trunc if variant.startswith('lt') else trunc + 1
        for n in range(1, hi):
            M2 = sum(pw[p] @ QL @ pw[n - p - 1] for p in range(n))
            m += M2 * T ** n / factorial(n)
        P = H.expm_rev(Q, pi, T)
        if variant.endswith('nodiv'):
            return m / P  # placeholder
        return m / P
    return counts

for variant, trunc in (('lt', 10), ('le', 10), ('lt', 9), ('lt', 11), ('lt', 20)):
    H.unif_counts_np = make(variant, trunc)
    r = H.map_np(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], c["codes"], c["code_mask"])
    n = r["n"].T if r["n"].shape != G.shape else r["n"]
    d = np.abs(n - G)
    print(variant, trunc, "max abs", d.max(), "median", np.median(d), "rel", (d / np.maximum(np.abs(G), 1e-3)).max())

H.unif_counts_np = make('lt', 10)
r = H.map_np(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], c["codes"], c["code_mask"])
n = r["n"].T if r["n"].shape != G.shape else r["n"]
d = np.abs(n - G).max(1)
order = np.argsort(-d)[:12]
print("rates", c["rates"])
for b in order:
    s = np.argmax(np.abs(n[b] - G[b]))
    print(b, "len", mean[b], "maxdiff", d[b], "mine", n[b, s], "gold", G[b, s], "unif", U[b, s])

for trunc in (3, 4, 5, 6, 7, 8):
    H.unif_counts_np = make('lt', trunc)
    r = H.map_np(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], c["codes"], c["code_mask"])
    n = r["n"].T if r["n"].shape != G.shape else r["n"]
    d = np.abs(n - G)
    print('lt', trunc, "max abs", d.max(), "b5", n[5, np.argmax(np.abs(G[5]-U[5]))], "median", np.median(d))
