"""Aggregates an `ncu --page source --csv --print-source sass` export: executed instructions and stall samples
per opcode, plus the hottest instructions.   python tools/ncu_src.py <src.csv> [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ex = collections.Counter(); st = collections.Counter(); tot_ex = 0; tot_s = 0
hot = []
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stall_tot = collections.Counter()
for k, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    e = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    ex[op] += e; st[op] += s; tot_ex += e; tot_s += s
    for c in stall_cols: stall_tot[c] += int(r[ix[c]] or 0)
    hot.append((s, e, k, src))
print("total executed %d, samples %d" % (tot_ex, tot_s))
print("%-10s %12s %6s %8s %6s" % ("opcode", "executed", "%", "samples", "%"))
for op, e in ex.most_common(28):
    print("%-10s %12d %6.1f %8d %6.1f" % (op, e, 100. * e / tot_ex, st[op], 100. * st[op] / max(1, tot_s)))
print("stall reasons:", ", ".join("%s %.1f%%" % (c[6:], 100. * v / max(1, tot_s)) for c, v in stall_tot.most_common(10)))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("hottest instructions (samples, executed, index, sass):")
for s, e, k, src in sorted(hot, reverse=True)[:top]:
    print("%6d %10d %5d  %s" % (s, e, k, src))
