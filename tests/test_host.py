"""Host front-end (comap_b200/host, C++): option syntax, readers, site selection, models and
rate distributions, checked through `comap_b200 --dry-run` (no GPU needed) against the
test-side Python readers, scipy and the reference's golden header."""
import os
import subprocess
import numpy as np
import pytest
import helpers as H
from comap_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "comap_b200", "bin", "comap_b200")


@pytest.fixture(scope="module")
def binary():
    from comap_b200 import build as b
    b.build()
    b.build_host()
    assert os.path.exists(BIN)
    return BIN


def dry_run(binary, cwd, *args):
    p = subprocess.run([binary] + list(args) + ["--dry-run"], cwd=cwd, capture_output=True, text=True)
    out = {}
    for ln in p.stdout.split("\n"):
        if ln.startswith("DRYRUN "):
            parts = ln.split()
            out[parts[1]] = parts[2:]
    return p, out


def write_fixture(tmp, name):
    g = H.golden(name)
    files = {"myoglobin": {"Myoglobin.aln.sel.mase": "mase", "Myo.dnd": "dnd", "comap.bpp": "options"},
             "bacteria_ssu": {"Bacteria_SSU.40.phy": "phy", "Bacteria_SSU.ML.dnd": "dnd", "options.comap": "options",
                              "options_pbs.mica": "mica_pbs", "options_npbs.mica": "mica_npbs", "options_perm.mica": "mica_perm",
                              "options_zscore.mica": "mica_zscore"},
             "srk": {"SRK.mase": "mase", "SRK.dnd": "dnd"}}[name]
    for fn, key in files.items():
        with open(os.path.join(tmp, fn), "wb") as f:
            f.write(bytes(g[key]))
    return g


def decode(out):
    n = int(out["n_nodes"][0])
    A = int(out["A"][0])
    d = dict(parent=np.array(out["parent"], dtype=np.int64), brlen=np.array(out["brlen"], dtype=float),
             leaves=out["leaves"], Q=np.array(out["Q"], dtype=float).reshape(A, A), pi=np.array(out["pi"], dtype=float),
             rates=np.array(out["rates"], dtype=float), probs=np.array(out["probs"], dtype=float),
             coords=np.array(out["coords"], dtype=np.int64), code_mask=np.array(out["code_mask"], dtype=np.uint32))
    T = len(d["leaves"])
    d["codes"] = np.array(out["codes"], dtype=np.uint8).reshape(T, -1)
    assert len(d["parent"]) == n
    return d


def test_myoglobin_inputs_match_python_readers_and_golden_header(binary, tmp_path):
    g = write_fixture(str(tmp_path), "myoglobin")
    p, out = dry_run(binary, str(tmp_path), "param=comap.bpp", "input.sequence.file=Myoglobin.aln.sel.mase",
                     "input.tree.file=Myo.dnd")
    assert p.returncode == 0, p.stdout
    d = decode(out)
    ref = H.myoglobin_inputs()
    assert np.array_equal(d["coords"], g["vec_coords"])           # the reference's own .vec header
    assert np.array_equal(d["parent"], ref["parent"])
    assert np.allclose(d["brlen"], ref["brlen"], rtol=0, atol=0)
    assert np.allclose(d["Q"], ref["Q"], rtol=1e-13, atol=1e-15)
    assert np.allclose(d["pi"], ref["pi"], rtol=1e-14)
    assert np.allclose(d["rates"], ref["rates"], rtol=1e-11)       # own incomplete gamma vs scipy
    assert np.allclose(d["probs"], ref["probs"])
    # same state sets cell by cell (code numbering is the binder's business)
    assert np.array_equal(d["code_mask"][d["codes"]], ref["code_mask"][ref["codes"]])
    assert "Number of sites to analyse.............: 129" in p.stdout


def test_bacteria_rna_gtr_invariant(binary, tmp_path):
    write_fixture(str(tmp_path), "bacteria_ssu")
    p, out = dry_run(binary, str(tmp_path), "param=options.comap")
    assert p.returncode == 0, p.stdout
    d = decode(out)
    assert d["codes"].shape == (40, 760)                           # 760 complete, variable sites
    th, th1, th2 = 0.523619444641, 0.512962941602, 0.585047306118
    pi = [th1 * (1 - th), (1 - th2) * th, th2 * th, (1 - th1) * (1 - th)]
    Q, pi = syn.gtr(1.595119085705, 0.551507085060, 0.350972557796, 0.304670173544, 0.282819006597, pi)
    assert np.allclose(d["Q"], Q, rtol=1e-13) and np.allclose(d["pi"], pi, rtol=1e-14)
    r, q = syn.gamma_rates(0.737023854405, 4)
    r, q = syn.invariant(r, q, 0.366611781033)
    assert np.allclose(d["rates"], r, rtol=1e-11) and np.allclose(d["probs"], q, rtol=1e-14)
    names, seqs = H.read_phylip_sequential_extended(H.text(H.golden("bacteria_ssu")["phy"]))
    parent, brlen, leaf_names = H.parse_newick(H.text(H.golden("bacteria_ssu")["dnd"]))
    assert d["leaves"] == leaf_names and np.array_equal(d["parent"], parent)
    assert np.array_equal(d["brlen"], brlen)
    row = {n: i for i, n in enumerate(names)}
    for k, leaf in enumerate(leaf_names):                          # resolved characters only ("complete")
        s = seqs[row[leaf]]
        assert "".join("ACGU"[c] for c in d["codes"][k]) == "".join(s[c - 1] for c in d["coords"])


def test_mica_option_files_parse_as_shipped(binary, tmp_path):
    """examples/RNA/BacteriaSSU/options_{pbs,npbs,perm,zscore}.mica: mica's option files select the same 760 sites as
    CoMap's; the parametric bootstrap's file carries the GTR + Invariant(Gamma4) model, the others none; the dry run lists
    the sequences in the alignment's own order (mica without a model keeps it)."""
    write_fixture(str(tmp_path), "bacteria_ssu")
    names, _ = H.read_phylip_sequential_extended(H.text(H.golden("bacteria_ssu")["phy"]))
    ref = decode(dry_run(binary, str(tmp_path), "param=options.comap")[1])
    for opt in ("options_pbs.mica", "options_npbs.mica", "options_perm.mica", "options_zscore.mica"):
        p, out = dry_run(binary, str(tmp_path), "param=" + opt)
        assert p.returncode == 0, p.stdout
        d = decode(out)
        assert d["codes"].shape == (40, 760) and np.array_equal(d["coords"], ref["coords"])
        assert np.array_equal(d["codes"], ref["codes"]) and out["sequences"] == names
        if opt == "options_pbs.mica":
            assert np.allclose(d["Q"], ref["Q"]) and np.allclose(d["rates"], ref["rates"]) and len(d["rates"]) == 5


def test_empirical_model_from_a_paml_file(binary, tmp_path):
    """model = Empirical(file=...) (Bio++ syntax) reads a PAML exchangeability file: the bundled JTT92
    table through that route equals model = JTT92."""
    write_fixture(str(tmp_path), "myoglobin")
    dat = os.path.join(os.path.dirname(os.path.dirname(binary)), "data", "jtt92_dcmut.dat")
    common = ["param=comap.bpp", "input.sequence.file=Myoglobin.aln.sel.mase", "input.tree.file=Myo.dnd"]
    p1, o1 = dry_run(binary, str(tmp_path), *common, "model=JTT92")
    p2, o2 = dry_run(binary, str(tmp_path), *common, "model=Empirical(name=myJTT, file=%s)" % dat)
    assert p1.returncode == 0 and p2.returncode == 0, p2.stdout
    assert o1["Q"] == o2["Q"] and o1["pi"] == o2["pi"]
    # LG08 (what examples/simple/* ask for) is bundled but unverified (re-typed from memory): it runs with a warning
    p3, o3 = dry_run(binary, str(tmp_path), *common, "model=LG08")
    assert p3.returncode == 0 and "NOT validated" in p3.stdout and "lg08.dat" in p3.stdout
    Q3, pi3 = np.array(o3["Q"], float).reshape(20, 20), np.array(o3["pi"], float)
    assert abs(pi3.sum() - 1) < 1e-12 and np.allclose(Q3.sum(1), 0, atol=1e-12) and abs((pi3 * np.diag(Q3)).sum() + 1) < 1e-12
    assert np.allclose(pi3[:, None] * Q3, (pi3[:, None] * Q3).T, atol=1e-15) and o3["Q"] != o1["Q"]
    p4, _ = dry_run(binary, str(tmp_path), *common, "model=WAG01")
    assert p4.returncode == 255 and "wag01.dat" in p4.stdout


def test_count_weights_and_second_data_set_options(binary, tmp_path):
    """weight=Diff(index1=Volume, symmetrical=no) (examples/simple/ProteinPairCompensation/comap.bpp:48) and the
    KEY2 options of a second data set (CoETools::readData suffix, CoETools.cpp:91-93; CoMap.cpp:236-262)."""
    write_fixture(str(tmp_path), "myoglobin")
    common = ["param=comap.bpp", "input.sequence.file=Myoglobin.aln.sel.mase", "input.tree.file=Myo.dnd"]
    p, o = dry_run(binary, str(tmp_path), *common, "nijt=Uniformization(weight=Diff(index1=Volume, symmetrical=no))")
    assert p.returncode == 0, p.stdout
    W = np.array(o["weights"], dtype=float).reshape(20, 20)
    assert W[0, 1] == 124 - 31 and W[1, 0] == 31 - 124 and np.array_equal(W, -W.T)   # A -> R gains 93 A^3 (Grantham 1974)
    p, o = dry_run(binary, str(tmp_path), *common, "nijt=Decomposition(weight=Diff(index1=Charge, symmetrical=yes))")
    W = np.array(o["weights"], dtype=float).reshape(20, 20)
    assert np.array_equal(W, W.T) and W[1, 3] == 2 and o["count_method"] == ["1"]        # R(+1) <-> D(-1)
    p, o = dry_run(binary, str(tmp_path), *common, "nijt=Naive(weight=AAdist(type=grantham, sym=yes))")
    W = np.array(o["weights"], dtype=float).reshape(20, 20)                              # Grantham 1974
    assert o["count_method"] == ["2"] and W[0, 1] == 112 and W[4, 17] == 215 and W[9, 10] == 5 and np.array_equal(W, W.T)
    p, o = dry_run(binary, str(tmp_path), *common)
    assert o["weights"] == []
    # nijt=Laplace (analyse.sh:12-15): the truncation order rides in the count word, 10 by default; no weights
    p, o = dry_run(binary, str(tmp_path), *common, "nijt=Laplace")
    assert p.returncode == 0 and o["count_method"] == [str(3 | 10 << 8)]
    p, o = dry_run(binary, str(tmp_path), *common, "nijt=Laplace(trunc=6)")
    assert o["count_method"] == [str(3 | 6 << 8)]
    p, o = dry_run(binary, str(tmp_path), *common, "nijt=ProbOneJump")
    assert p.returncode == 0 and o["count_method"] == ["5"]
    p, _ = dry_run(binary, str(tmp_path), *common, "nijt=Laplace(weight=Diff(index1=Volume, symmetrical=no))")
    assert p.returncode == 255 and "does not take weights" in p.stdout
    # second data set: same files, all sites instead of the complete ones; the tree is copied
    p = subprocess.run([binary] + common + ["input.sequence.file2=Myoglobin.aln.sel.mase", "input.sequence.sites_to_use2=all",
                                            "--dry-run"], cwd=str(tmp_path), capture_output=True, text=True)
    assert p.returncode == 0, p.stdout
    first, second = p.stdout.split("DRYRUN second_data_set 1")
    get = lambda txt, key: [ln.split()[2:] for ln in txt.split("\n") if ln.startswith("DRYRUN " + key + " ")][0]
    assert get(first, "parent") == get(second, "parent") and get(first, "Q") == get(second, "Q")
    assert len(get(second, "coords")) > len(get(first, "coords")) == 129


def test_mase_site_selection_srk(binary, tmp_path):
    write_fixture(str(tmp_path), "srk")
    p, out = dry_run(binary, str(tmp_path), "alphabet=Protein", "input.sequence.file=SRK.mase",
                     "input.sequence.format=Mase(site_selection=SelectedSites)", "input.sequence.sites_to_use=all",
                     "input.remove_const=no", "input.tree.file=SRK.dnd", "model=JTT92", "rate_distribution=Constant()")
    assert p.returncode == 0, p.stdout
    d = decode(out)
    segs = [(23, 31), (35, 66), (70, 82), (86, 150), (156, 209), (215, 221), (227, 307), (312, 433), (440, 452)]
    want = [c for a, b in segs for c in range(a, b + 1)]
    assert d["coords"].tolist() == want
    assert d["rates"].tolist() == [1.0]


def _tiny(tmp, fasta, tree, opts):
    with open(os.path.join(tmp, "a.fa"), "w") as f:
        f.write(fasta)
    with open(os.path.join(tmp, "t.dnd"), "w") as f:
        f.write(tree)
    with open(os.path.join(tmp, "o.bpp"), "w") as f:
        f.write(opts)


FASTA = ">s1 first\nACGTNA-C\n>s2\nACGTAAGC\n>s3\nACCTAARC\n>s4\nATCTAAGC\n"
TREE = "((s1:0.1,s2:0.2)0.9:0.05,s3:0.3,s4:1e-9);\n"
OPTS = """# comment line
DATA = a            # trailing comment
alphabet = DNA
input.sequence.file = $(DATA).fa
input.sequence.format = Fasta
input.sequence.sites_to_use = all
input.remove_const = no
input.tree.file = t.dnd
model = HKY85(kappa = 2.5, theta=0.4, \\
              theta1 = 0.6, theta2=0.5)
rate_distribution = Gamma(n=3, alpha=0.7)
"""


def test_option_syntax_and_overrides(binary, tmp_path):
    _tiny(str(tmp_path), FASTA, TREE, OPTS)
    p, out = dry_run(binary, str(tmp_path), "param=o.bpp")
    assert p.returncode == 0, p.stdout
    d = decode(out)
    Q, pi = syn.hky85(2.5, [0.6 * 0.6, 0.5 * 0.4, 0.5 * 0.4, 0.4 * 0.6])
    assert np.allclose(d["Q"], Q, rtol=1e-13) and np.allclose(d["pi"], pi)
    r, q = syn.gamma_rates(0.7, 3)
    assert np.allclose(d["rates"], r, rtol=1e-11)
    assert d["leaves"] == ["s1", "s2", "s3", "s4"] and d["parent"].tolist() == [2, 2, 5, 5, 5, -1]
    assert d["brlen"].tolist() == [0.1, 0.2, 0.05, 0.3, 1e-9, 0.0]   # the 1e-6 floor is applied by the library
    assert d["coords"].tolist() == list(range(1, 9))
    # N, '-' -> any state; R -> A|G
    m = d["code_mask"][d["codes"]]
    assert m[0].tolist() == [1, 2, 4, 8, 15, 1, 15, 2] and m[2, 6] == 5
    # command line wins; sites_to_use / remove_const
    p, out = dry_run(binary, str(tmp_path), "param=o.bpp", "input.sequence.sites_to_use=nogap", "input.remove_const=yes",
                     "model=K80(kappa=3)", "rate_distribution=Invariant(dist=Gamma(n=2,alpha=1.0),p=0.25)")
    d = decode(out)
    assert d["coords"].tolist() == [2, 3]                            # cols 1,4,5,6,8 constant; 7 has a gap
    assert np.allclose(d["pi"], 0.25) and np.isclose(d["Q"][0, 2] / d["Q"][0, 1], 3.0)
    r, q = syn.invariant(*syn.gamma_rates(1.0, 2), 0.25)
    assert np.allclose(d["rates"], r, rtol=1e-11) and np.allclose(d["probs"], q)
    p, out = dry_run(binary, str(tmp_path), "param=o.bpp", "input.sequence.sites_to_use=complete")
    assert decode(out)["coords"].tolist() == [1, 2, 3, 4, 6, 8]


def test_rooted_tree_is_unrooted_and_errors_are_reported(binary, tmp_path):
    _tiny(str(tmp_path), FASTA, "((s1:0.1,s2:0.2):0.05,(s3:0.3,s4:0.4):0.07);", OPTS)
    p, out = dry_run(binary, str(tmp_path), "param=o.bpp")
    assert p.returncode == 0 and "Tree has been unrooted" in p.stdout
    d = decode(out)
    assert len(d["parent"]) == 6 and (d["parent"] == -1).sum() == 1 and np.isclose(d["brlen"].sum(), 0.1 + 0.2 + 0.3 + 0.4 + 0.12)
    for bad, msg in (("model=LG08", "not supported for nucleotides"), ("nijt=Bogus", "not available"),
                     ("input.tree.file=missing.dnd", "cannot open"), ("alphabet=Codon", "not supported"),
                     ("rate_distribution=Gamma(n=0,alpha=1)", "n must be positive")):
        p, out = dry_run(binary, str(tmp_path), "param=o.bpp", bad)
        assert p.returncode != 0 and msg in p.stdout, (bad, p.stdout[-300:])
    _tiny(str(tmp_path), FASTA.replace(">s4", ">zz"), TREE, OPTS)
    p, out = dry_run(binary, str(tmp_path), "param=o.bpp")
    assert p.returncode != 0 and "has no sequence" in p.stdout


def test_phylip_interleaved_and_classic(binary, tmp_path):
    phy = " 4 8\ns1        ACGT\ns2        ACGT\ns3        ACCT\ns4        ATCT\n\nNA-C\nAAGC\nAARC\nAAGC\n"
    _tiny(str(tmp_path), FASTA, TREE, OPTS)
    with open(os.path.join(str(tmp_path), "a.phy"), "w") as f:
        f.write(phy)
    p1, o1 = dry_run(binary, str(tmp_path), "param=o.bpp")
    p2, o2 = dry_run(binary, str(tmp_path), "param=o.bpp", "input.sequence.file=a.phy",
                     "input.sequence.format=Phylip(order=interleaved, type=classic)")
    assert p2.returncode == 0, p2.stdout
    assert o1["codes"] == o2["codes"] and o1["code_mask"] == o2["code_mask"]
