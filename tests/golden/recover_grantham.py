"""Recovers Bio++'s Grantham chemical-distance table (weight=AAdist(type=grantham, sym=yes)) from the
reference's own golden output (run in the build container only; needs /root/reference).

examples/Proteins/Benchmark/CoMap/Myo_naive_grantham.vec was written by
`nijt=Naive(weight=AAdist(type=grantham, sym=yes))` (analyse.sh): with the naive count every
mapping entry is  sum over amino-acid pairs {x,y} of  w(x,y) * c_xy(site, branch),  c_xy = the
posterior weight of an x<->y change, which the oracle yields by mapping with an indicator weight
matrix.  25 413 equations, 190 unknowns: least squares, then entries within 0.12 of an integer are
fixed and the rest re-solved.  187 entries come out as integers at once; N-W, G-W and P-W (hardly
ever observed in Myoglobin) round to 174, 184, 147.  The result equals Grantham's (1974) published
table and is written to comap_b200/data/grantham.dat; tests/test_oracle_golden.py then checks the
weighted Uniformization / Decomposition / Naive mappings against the three *_grantham.vec files.
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, HERE)
import helpers as H, oracle_binding as O
from make_golden import load_vec

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
coords, mean, G = load_vec(REF + "/examples/Proteins/Benchmark/CoMap/Myo_naive_grantham.vec")
m = H.myoglobin_inputs()
gold = G.T.ravel()
pairs = [(x, y) for x in range(20) for y in range(x + 1, 20)]
cols = []
for x, y in pairs:
    W = np.zeros((20, 20)); W[x, y] = W[y, x] = 1
    r = O.map_sites(m["parent"], m["brlen"], m["Q"], m["pi"], m["rates"], m["probs"], m["codes"], m["code_mask"],
                    method="naive", weights=W)
    cols.append(r["n"].ravel())
Cm = np.stack(cols, axis=1)
w = np.linalg.lstsq(Cm, gold, rcond=None)[0]
fixed = {}
for it in range(8):
    for k in range(190):
        if k not in fixed and abs(w[k] - round(w[k])) < 0.12:
            fixed[k] = round(w[k])
    free = [k for k in range(190) if k not in fixed]
    if not free:
        break
    rhs = gold - Cm[:, list(fixed)] @ np.array([fixed[k] for k in fixed], float)
    sol = np.linalg.lstsq(Cm[:, free], rhs, rcond=None)[0]
    for k, v in zip(free, sol):
        w[k] = v
print("entries left to rounding:", [("ARNDCQEGHILKMFPSTWYV"[pairs[k][0]] + "ARNDCQEGHILKMFPSTWYV"[pairs[k][1]], round(float(w[k]), 3))
                                     for k in range(190) if k not in fixed])
D = np.zeros((20, 20), dtype=int)
for k, (x, y) in enumerate(pairs):
    D[x, y] = D[y, x] = fixed.get(k, int(round(w[k])))
out = os.path.join(os.path.dirname(os.path.dirname(HERE)), "comap_b200", "data", "grantham.dat")
with open(out, "w") as f:
    f.write("# Grantham (1974) chemical distance, order A R N D C Q E G H I L K M F P S T W Y V (Bio++ protein alphabet);\n"
            "# recovered from the reference's Myo_naive_grantham.vec by tests/golden/recover_grantham.py\n")
    for row in D:
        f.write(" ".join("%d" % v for v in row) + "\n")
print("wrote", out, "max", D.max(), "C-W", D[4, 17], "I-L", D[9, 10])
