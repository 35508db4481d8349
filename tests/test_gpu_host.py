"""End-to-end through the C++ front-end (`comap_b200 param=...`): the reference's option
file for the Myoglobin example, output tables compared with the reference's own golden
files (mapping, infos) and with the same analysis driven through the C ABI from Python."""
import os
import subprocess
import numpy as np
import pytest
import helpers as H
import oracle_binding as O
from test_host import BIN, write_fixture, dry_run, decode

pytestmark = pytest.mark.gpu


def run(cwd, *args):
    p = subprocess.run([BIN] + list(args), cwd=cwd, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


def table(path):
    rows = [ln.rstrip("\n").split("\t") for ln in open(path)]
    return rows[0], rows[1:]


def g(v):
    return "%g" % v


@pytest.fixture(scope="module")
def myo(tmp_path_factory):
    from comap_b200 import build as b
    b.build_host()
    tmp = str(tmp_path_factory.mktemp("myo"))
    golden = write_fixture(tmp, "myoglobin")
    return tmp, golden


def host_inputs(tmp):
    """Exactly the arrays the binary hands to the library (17-digit dump): parameters computed
    by the C++ host (own incomplete gamma, own normalisation) differ from numpy/scipy in the last
    bits, which is enough to break ties differently in clustering."""
    p, out = dry_run(BIN, tmp, *COMMON[:-1])
    assert p.returncode == 0
    d = decode(out)
    d["parent"] = d["parent"].astype(np.int32)
    return d


COMMON = ["param=comap.bpp", "input.sequence.file=Myoglobin.aln.sel.mase", "input.tree.file=Myo.dnd",
          "nijt=Uniformization", "--seed=11"]


def test_laplace_count_from_the_command_line_matches_its_golden(myo):
    """examples/Proteins/Benchmark/CoMap/analyse.sh:12-15: comap param=comap.bpp nijt=Laplace -> Myo_laplace.vec."""
    tmp, golden = myo
    run(tmp, *[a for a in COMMON if not a.startswith("nijt=")], "nijt=Laplace", "analysis=none", "output.vectors.file=Myo_laplace.vec")
    _, rows = table(os.path.join(tmp, "Myo_laplace.vec"))
    vec = np.array([[float(x) for x in r[2:]] for r in rows])
    assert vec.shape == (197, 129)
    err = np.abs(vec - golden["vec_laplace"])
    assert err.max() < 3e-5 and np.median(err / np.abs(golden["vec_laplace"])) < 1e-5


def test_mapping_only_matches_reference_golden_files(myo):
    tmp, golden = myo
    run(tmp, *COMMON, "analysis=none", "output.vectors.file=Myo.vec", "output.infos=Myo.infos")
    hdr, rows = table(os.path.join(tmp, "Myo.vec"))
    assert hdr[:2] == ["Branches", "Mean"] and hdr[2:] == ["Site%d" % c for c in golden["vec_coords"]]
    vec = np.array([[float(x) for x in r[2:]] for r in rows])
    assert vec.shape == (197, 129) and [int(r[0]) for r in rows] == list(range(197))
    assert np.allclose([float(r[1]) for r in rows], golden["vec_brlen"], rtol=1e-5)
    # both files hold 6 significant digits; same bar as tests/test_oracle_golden.py (the golden was
    # produced in 2012 by an older Bio++), widened by one printed ulp
    rel = np.abs(vec - golden["vec_unif"]) / np.abs(golden["vec_unif"])
    assert np.median(rel) < 1e-5 and rel.max() < 1.2e-4 and (rel > 2e-5).mean() < 0.1
    hdr, rows = table(os.path.join(tmp, "Myo.infos"))
    assert hdr == ["Group", "IsComplete", "IsConstant", "RC", "PR", "N", "logLn"]   # CoETools.cpp:515
    assert [r[0] for r in rows] == ["[%d]" % c for c in golden["infos_coord"]]
    assert [int(r[1]) for r in rows] == golden["infos_complete"].tolist()
    assert [int(r[2]) for r in rows] == golden["infos_const"].tolist()
    assert [int(r[3]) for r in rows] == golden["infos_rc"].tolist()
    assert np.allclose([float(r[4]) for r in rows], golden["infos_pr"], rtol=2e-5)
    assert np.allclose([float(r[6]) for r in rows], golden["infos_logl"], rtol=2e-5)


def test_pairwise_table_equals_c_abi_run(myo):
    from comap_b200 import api
    tmp, _ = myo
    run(tmp, *COMMON, "analysis=pairwise", "statistic=Correlation", "statistic.output.file=stats.txt",
        "statistic.null.nb_rep_CPU=3", "statistic.null.nb_rep_RAM=200", "statistic.null.nb_rate_classes=4",
        "statistic.null.output.file=null.txt")
    c = host_inputs(tmp)
    ctx = api.Context(device=0)
    ctx.set_tree(c["parent"], c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
    ctx.set_alignment(c["codes"], c["code_mask"])
    m = ctx.map()
    raw = ctx.null_intra("correlation", 11, 3, 200, K=4, want_raw=True)
    p, k = ctx.pairs("correlation", use_null=True)
    hdr, rows = table(os.path.join(tmp, "stats.txt"))
    assert hdr == ["Group", "Stat", "RCmin", "PRmin", "Nmin", "PValue", "Nsim"]      # CoETools.cpp:662-665
    assert len(rows) == k == 129 * 128 // 2
    co = c["coords"]
    for r in (0, 1, 127, 128, 5000, k - 1):
        want = ["[%d;%d]" % (co[p["i"][r]], co[p["j"][r]]), g(p["stat"][r]), str(p["rcmin"][r]), g(p["prmin"][r]),
                g(p["nmin"][r])] + (["NA", "0"] if np.isnan(p["pvalue"][r]) else [g(p["pvalue"][r]), str(p["nsim"][r])])
        assert rows[r] == want
    assert [x[5] for x in rows] == ["NA" if np.isnan(v) else g(v) for v in p["pvalue"]]
    assert [x[1] for x in rows] == [g(v) for v in p["stat"]]
    # statistics against the oracle on the oracle's own mapping (6 printed digits)
    q = O.map_sites(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], c["codes"], c["code_mask"])
    op = O.pairs("correlation", q["n"], q["norm"], q["post_rate"], q["rate_class"])
    assert np.allclose([float(x[1]) for x in rows], op["stat"], rtol=2e-5, atol=2e-6)
    hdr, nrows = table(os.path.join(tmp, "null.txt"))
    assert hdr == ["Stat", "RCmin", "PRmin", "Nmin"] and len(nrows) == 600            # AnalysisTools.cpp:580,642
    assert [x[0] for x in nrows] == [g(v) for v in raw[:, 0]] and [x[3] for x in nrows] == [g(v) for v in raw[:, 3]]
    ctx.close()


def test_clustering_tables_equal_c_abi_run(myo):
    from comap_b200 import api
    tmp, _ = myo
    run(tmp, *COMMON, "analysis=clustering", "clustering.distance=cor", "clustering.method=complete",
        "clustering.output.groups.file=groups.txt", "clustering.output.matrix.file=mat.phy",
        "clustering.output.tree.file=clust.dnd", "clustering.null=yes", "clustering.null.number=2",
        "clustering.null.output.file=groups_null.txt", "clustering.maximum_group_size=6")
    c = host_inputs(tmp)
    ctx = api.Context(device=0)
    ctx.set_tree(c["parent"], c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
    ctx.set_alignment(c["codes"], c["code_mask"])
    ctx.map()
    mat = ctx.distance_matrix("correlation")
    ctx.cluster("complete")
    grp = ctx.groups("correlation", 6)
    hdr, rows = table(os.path.join(tmp, "groups.txt"))
    assert hdr == ["Group", "Size", "IsConstant", "Dmax", "Stat", "Nmin"]              # CoMap.cpp:494-550
    assert len(rows) == len(grp["members"])
    co = c["coords"]
    for r, mem in zip(rows, grp["members"]):
        assert r[0] == "[" + ";".join(str(co[x]) for x in mem) + "]" and int(r[1]) == len(mem) and r[2] == "no"
    assert [r[3] for r in rows] == [g(2 * h) for h in grp["height"]]
    assert [r[4] for r in rows] == [g(v) for v in grp["stat"]] and [r[5] for r in rows] == [g(v) for v in grp["nmin"]]
    lines = open(os.path.join(tmp, "mat.phy")).read().split("\n")
    assert int(lines[0]) == 129
    first = lines[1].split()
    assert first[0] == str(co[0]) and [float(x) for x in first[1:]] == [float(g(v)) for v in mat[0]]
    nul = ctx.cluster_null("correlation", "complete", 11, 0, 2, 6)
    hdr, rows = table(os.path.join(tmp, "groups_null.txt"))
    assert hdr == ["Rep", "Group", "Size", "Dmax", "Stat", "Nmin"]                     # ClusterTools.cpp:219
    assert [int(r[0]) for r in rows] == nul["rep"].tolist()
    assert [r[1] for r in rows] == ["[" + ";".join(str(x) for x in mem) + "]" for mem in nul["members"]]
    assert [r[3] for r in rows] == [g(v) for v in nul["dmax"]] and [r[4] for r in rows] == [g(v) for v in nul["stat"]]
    tree = open(os.path.join(tmp, "clust.dnd")).read().strip()
    assert tree.endswith(";") and tree.count("(") == 128 and tree.count(",") == 128
    parent, brlen, leaf_names = H.parse_newick(tree)
    assert sorted(int(x) for x in leaf_names) == sorted(co.tolist())
    ctx.close()


def test_restart_from_the_reference_mapping_file(myo):
    """input.vectors.file (CoETools.cpp:374-385): the pairwise analysis restarted from the
    REFERENCE's own golden mapping file equals the statistics of those vectors."""
    tmp, golden = myo
    with open(os.path.join(tmp, "golden.vec"), "w") as f:
        f.write("Branches\tMean" + "".join("\tSite%d" % c for c in golden["vec_coords"]) + "\n")
        for b in range(197):
            f.write("%d\t%g" % (b, golden["vec_brlen"][b]) + "".join("\t%.6g" % v for v in golden["vec_unif"][b]) + "\n")
    run(tmp, *COMMON, "analysis=pairwise", "statistic=Correlation", "input.vectors.file=golden.vec",
        "statistic.null=no", "statistic.output.file=restart.txt", "output.infos=restart.infos")
    hdr, rows = table(os.path.join(tmp, "restart.txt"))
    assert hdr == ["Group", "Stat", "RCmin", "PRmin", "Nmin"] and len(rows) == 129 * 128 // 2
    vec = np.array([[float("%.6g" % v) for v in row] for row in golden["vec_unif"]]).T   # [site][branch]
    for r in (0, 1, 500, 8000, len(rows) - 1):
        i, j = [list(golden["vec_coords"]).index(int(x)) for x in rows[r][0].strip("[]").split(";")]
        assert rows[r][1] == g(O.stat("correlation", vec[i], vec[j]))
        assert rows[r][4] == g(min(np.sqrt((vec[i] ** 2).sum()), np.sqrt((vec[j] ** 2).sum())))
    hdr, irows = table(os.path.join(tmp, "restart.infos"))                                 # N column = loaded norms
    assert np.allclose([float(r[5]) for r in irows], np.sqrt((vec ** 2).sum(axis=1)), rtol=2e-5)


def test_two_data_sets_table_equals_c_abi_run(myo):
    """input.sequence.file2 -> inter-gene flow (CoMap.cpp:236-347): rectangle of statistics without
    p-values, per-data-set rate thresholds (KEY2), null rows of the two simulators."""
    from comap_b200 import api
    tmp, _ = myo
    run(tmp, *COMMON, "analysis=pairwise", "statistic=Correlation", "input.sequence.file2=Myoglobin.aln.sel.mase",
        "statistic.output.file=inter.txt", "statistic.min_rate_class2=2", "statistic.min=0.2", "output.infos2=infos2.txt",
        "statistic.null.nb_rep_CPU=2", "statistic.null.nb_rep_RAM=150", "statistic.null.output.file=inter_null.txt")
    c = host_inputs(tmp)
    a, b = api.Context(device=0), api.Context(device=0)
    for ctx in (a, b):
        ctx.set_tree(c["parent"], c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
        ctx.set_alignment(c["codes"], c["code_mask"]); ctx.map()
    p, k = a.pairs_inter(b, "correlation", filters=dict(min_stat=0.2), min_rate_class2=2)
    hdr, rows = table(os.path.join(tmp, "inter.txt"))
    assert hdr == ["Group", "Stat", "RCmin", "PRmin", "Nmin"]                           # CoETools.cpp:779
    assert 0 < len(rows) == k < 129 * 129
    co = c["coords"]
    assert [x[0] for x in rows] == ["[%d;%d]" % (co[i], co[j]) for i, j in zip(p["i"], p["j"])]
    assert [x[1] for x in rows] == [g(v) for v in p["stat"]] and [x[4] for x in rows] == [g(v) for v in p["nmin"]]
    raw = a.null_inter(b, "correlation", 11, 2, 150)
    hdr, nrows = table(os.path.join(tmp, "inter_null.txt"))
    assert hdr == ["Stat", "RCmin", "PRmin", "Nmin"] and len(nrows) == 300
    assert [x[0] for x in nrows] == [g(v) for v in raw[:, 0]]
    hdr, irows = table(os.path.join(tmp, "infos2.txt"))
    assert hdr[0] == "Group" and len(irows) == 129
    a.close(); b.close()


VOLUME = np.array([31, 124, 56, 54, 55, 85, 83, 3, 96, 111, 111, 119, 105, 132, 32.5, 32, 61, 170, 136, 84.])  # Grantham 1974


def test_weighted_counts_and_compensation_statistic(myo):
    """nijt=Uniformization(weight=Diff(index1=Volume, symmetrical=no)) + statistic=Compensation
    (examples/simple/ProteinPairCompensation/comap.bpp:41-48; CoETools.cpp:564-574): signed
    volume-change mapping, statistic 1 - |v1+v2| / (|v1|+|v2|), against the oracle."""
    tmp, _ = myo
    args = [a for a in COMMON if not a.startswith("nijt=")]
    run(tmp, *args, "nijt=Uniformization(weight=Diff(index1=Volume, symmetrical=no))", "analysis=pairwise",
        "statistic=Compensation", "statistic.null=no", "statistic.output.file=comp.txt", "output.vectors.file=wvec.txt")
    c = host_inputs(tmp)
    W = VOLUME[None, :] - VOLUME[:, None]                       # w[x][y] = volume[y] - volume[x]
    q = O.map_sites(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], c["codes"], c["code_mask"], weights=W)
    hdr, rows = table(os.path.join(tmp, "wvec.txt"))
    vec = np.array([[float(x) for x in r[2:]] for r in rows]).T
    assert (vec < 0).any() and np.allclose(vec, q["n"], rtol=2e-5, atol=1e-9)      # signed counts, 6 printed digits
    op = O.pairs("compensation", q["n"], q["norm"], q["post_rate"], q["rate_class"])
    hdr, rows = table(os.path.join(tmp, "comp.txt"))
    assert len(rows) == len(op["i"]) == 129 * 128 // 2
    assert np.allclose([float(x[1]) for x in rows], op["stat"], rtol=2e-5, atol=2e-6)
    p = subprocess.run([BIN] + args + ["nijt=Uniformization(weight=Diff(index1=Volume, symmetrical=yes))", "analysis=pairwise",
                                       "statistic=Compensation"], cwd=tmp, capture_output=True, text=True)
    assert p.returncode == 255 and "non-symmetric weights" in p.stdout


def test_clustering_with_the_compensation_distance(myo):
    """examples/Proteins/GroupsCompensation: clustering.distance=comp on a signed volume-change mapping
    (CoMap.cpp:412-422; Distance.h:382-422): groups, Dmax and the compensation group statistic."""
    from comap_b200 import api
    tmp, _ = myo
    args = [a for a in COMMON if not a.startswith("nijt=")] + ["nijt=Uniformization(weight=Diff(index1=Volume, symmetrical=no))"]
    run(tmp, *args, "analysis=clustering", "clustering.distance=comp", "clustering.method=complete",
        "clustering.output.groups.file=cgroups.txt", "clustering.maximum_group_size=5", "clustering.null=no")
    c = host_inputs(tmp)
    W = VOLUME[None, :] - VOLUME[:, None]
    ctx = api.Context(device=0)
    ctx.set_tree(c["parent"], c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"], weights=W)
    ctx.set_alignment(c["codes"], c["code_mask"]); ctx.map()
    ctx.distance_matrix("compensation"); ctx.cluster("complete")
    grp = ctx.groups("compensation", 5)
    hdr, rows = table(os.path.join(tmp, "cgroups.txt"))
    assert hdr == ["Group", "Size", "IsConstant", "Dmax", "Stat", "Nmin"] and len(rows) == len(grp["members"]) > 10
    co = c["coords"]
    assert [r[0] for r in rows] == ["[" + ";".join(str(co[x]) for x in mem) + "]" for mem in grp["members"]]
    assert [r[4] for r in rows] == [g(v) for v in grp["stat"]] and [r[3] for r in rows] == [g(2 * h) for h in grp["height"]]
    p = subprocess.run([BIN] + COMMON + ["analysis=clustering", "clustering.distance=comp"], cwd=tmp, capture_output=True, text=True)
    assert p.returncode == 255 and "with weights" in p.stdout
    ctx.close()


def test_mutual_information_from_the_cli(myo):
    """statistic=MI(threshold=0.5) with nijt=Uniformization (CoETools.cpp:576-596)."""
    from comap_b200 import api
    tmp, _ = myo
    run(tmp, *COMMON, "analysis=pairwise", "statistic=MI(threshold=0.5)", "statistic.null=no", "statistic.output.file=mi.txt")
    c = host_inputs(tmp)
    ctx = api.Context(device=0)
    ctx.set_tree(c["parent"], c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
    ctx.set_alignment(c["codes"], c["code_mask"]); ctx.set_mi_threshold(0.5); ctx.map()
    p, k = ctx.pairs("mi", use_null=False)
    hdr, rows = table(os.path.join(tmp, "mi.txt"))
    assert len(rows) == k and [x[1] for x in rows] == [g(v) for v in p["stat"]]
    assert max(float(x[1]) for x in rows) > 0.05
    pr = subprocess.run([BIN] + [a for a in COMMON if not a.startswith("nijt=")] + ["nijt=Label", "analysis=pairwise", "statistic=MI"],
                        cwd=tmp, capture_output=True, text=True)
    assert pr.returncode == 255
    ctx.close()


def test_label_mutual_information_and_mapping_variants_from_the_cli(myo):
    """statistic=MI nijt=Label nijt.average=no (CoETools.cpp:577-589) and the nijt.average / nijt.joint switches of
    CoETools::getVectors (CoETools.cpp:393-407): the command line against the C ABI on the same inputs."""
    from comap_b200 import api
    tmp, _ = myo
    base = [a for a in COMMON if not a.startswith("nijt=")]
    run(tmp, *base, "nijt=Label", "nijt.average=no", "analysis=pairwise", "statistic=MI", "statistic.null=no",
        "statistic.output.file=mil.txt", "output.vectors.file=label.vec")
    c = host_inputs(tmp)
    ctx = api.Context(device=0)
    ctx.set_tree(c["parent"], c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"], count_method="label")
    ctx.set_alignment(c["codes"], c["code_mask"]); ctx.set_map_mode(False, True)
    r = ctx.map()
    p, k = ctx.pairs("mi_label", use_null=False)
    hdr, rows = table(os.path.join(tmp, "mil.txt"))
    assert len(rows) == k == 129 * 128 // 2 and [x[1] for x in rows] == [g(v) for v in p["stat"]]
    _, vrows = table(os.path.join(tmp, "label.vec"))
    vec = np.array([[float(x) for x in r_[2:]] for r_ in vrows])
    assert np.array_equal(vec, np.rint(r["n"].T)) and vec.max() > 100      # labels 0..380 of the 20-state alphabet
    # the three variants with the ordinary counts: vectors written by the CLI == C ABI
    for av, jo in (("yes", "no"), ("no", "yes"), ("no", "no")):
        run(tmp, *COMMON, "nijt.average=" + av, "nijt.joint=" + jo, "analysis=none", "output.vectors.file=v.vec")
        ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"]); ctx.set_alignment(c["codes"], c["code_mask"])
        ctx.set_map_mode(av == "yes", jo == "yes")
        n = ctx.map()["n"]
        _, vrows = table(os.path.join(tmp, "v.vec"))
        assert [x[2:] for x in vrows] == [[g(v) for v in row] for row in n.T]
    ctx.close()


def test_candidate_groups_table_equals_c_abi_run(myo):
    """analysis=candidates (CoMap.cpp:592-711): input table + Stat + p-value columns."""
    from comap_b200 import api
    tmp, _ = myo
    c = host_inputs(tmp)
    co = c["coords"]
    groups = [[3, 17], [40, 41, 90], [5, 100]]
    with open(os.path.join(tmp, "cand.txt"), "w") as f:
        f.write("Name\tGroup\n")
        for k, gsites in enumerate(groups):
            f.write("g%d\t[%s]\n" % (k, ";".join(str(co[x]) for x in gsites)))
        f.write("bad\t[%d;999999]\n" % co[0])                       # position that is not analysed -> NA
    run(tmp, *COMMON, "analysis=candidates", "statistic=Correlation", "candidates.input.file=cand.txt",
        "candidates.output.file=cand_out.txt", "candidates.null.min=30", "candidates.omega=1.0",
        "candidates.null.nb_rep_RAM=300", "candidates.nb_max_trials=3")
    ctx = api.Context(device=0)
    ctx.set_tree(c["parent"], c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
    ctx.set_alignment(c["codes"], c["code_mask"]); ctx.map()
    r = ctx.candidates("correlation", groups + [[0]], omega=1.0, min_sim=30, max_trials=3, rep_ram=300, seed=11,
                       analysable=[1, 1, 1, 0])
    hdr, rows = table(os.path.join(tmp, "cand_out.txt"))
    assert hdr == ["Name", "Group", "Stat", "p-value"] and len(rows) == 4
    for k in range(3):
        assert rows[k][2] == g(r["stat"][k]) and rows[k][3] == g(r["pvalue"][k])
    assert rows[3][2:] == ["NA", "NA"]
    assert r["n2"][:3].max() >= 1
    ctx.close()


def test_rna_example_with_its_own_option_file(tmp_path):
    """BASELINE configs[2]: examples/RNA/BacteriaSSU/options.comap unchanged -- GTR + Invariant(Gamma4)
    (five rate classes incl. a rate-0 class), 760 complete variable sites x 40 taxa, pairwise
    correlation with a 100 x 1000 null -- against the oracle on the inputs the binary used."""
    from comap_b200 import build as b
    b.build_host()
    tmp = str(tmp_path)
    write_fixture(tmp, "bacteria_ssu")
    out = run(tmp, "param=options.comap", "--seed=5")
    assert "Bye bye" in out
    hdr, rows = table(os.path.join(tmp, "Bacteria_SSU.sged"))
    assert hdr == ["Group", "Stat", "RCmin", "PRmin", "Nmin", "PValue", "Nsim"] and len(rows) == 760 * 759 // 2
    p, o = dry_run(BIN, tmp, "param=options.comap")
    c = decode(o); c["parent"] = c["parent"].astype(np.int32)
    q = O.map_sites(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], c["codes"], c["code_mask"])
    op = O.pairs("correlation", q["n"], q["norm"], q["post_rate"], q["rate_class"])
    st = np.array([float(x[1]) for x in rows])           # NaN rows: sites without any expected substitution
    ok = ~np.isnan(op["stat"])
    assert np.array_equal(np.isnan(st), ~ok)
    assert np.allclose(st[ok], op["stat"][ok], rtol=2e-5, atol=2e-6)
    assert [int(x[2]) for x in rows] == op["rcmin"].tolist()
    pv = np.array([float(x[5]) if x[5] != "NA" else np.nan for x in rows])
    assert np.nanmin(pv) > 0 and np.nanmax(pv) <= 1 and (~np.isnan(pv)).mean() > 0.5
    assert sum(int(x[6]) for x in rows[:2000]) > 0


def test_simple_examples_run_with_their_own_option_files(myo):
    """BASELINE configs[0] = examples/simple/ProteinPairCorrelation/comap.bpp as shipped (model = LG08, Gamma(4, 0.5),
    optimization = FullD -> parameters used as given, with a warning; 100 x 1000 null), and its siblings
    ProteinPairCompensation (weighted counts), ProteinGroupCorrelation / ProteinGroupCompensation (clustering, null cut
    to 3 replicates on the command line) and ProteinMappingOnly.  The LG08 table is bundled unverified (see
    comap_b200/data/lg08.dat), so the check is against the oracle on the inputs the binary used, not against Bio++."""
    tmp, golden = myo
    for ex in ("ProteinPairCorrelation", "ProteinPairCompensation", "ProteinGroupCorrelation", "ProteinGroupCompensation",
               "ProteinMappingOnly"):
        with open(os.path.join(tmp, ex + ".bpp"), "wb") as f:
            f.write(bytes(golden["simple_" + ex]))
    out = run(tmp, "param=ProteinPairCorrelation.bpp", "--seed=3")
    assert "NOT validated" in out and "outside the B200 hot path" in out and "Bye bye" in out
    hdr, rows = table(os.path.join(tmp, "Myo.results.txt"))
    assert hdr == ["Group", "Stat", "RCmin", "PRmin", "Nmin", "PValue", "Nsim"] and len(rows) == 129 * 128 // 2
    p, o = dry_run(BIN, tmp, "param=ProteinPairCorrelation.bpp")
    c = decode(o); c["parent"] = c["parent"].astype(np.int32)
    q = O.map_sites(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], c["codes"], c["code_mask"])
    op = O.pairs("correlation", q["n"], q["norm"], q["post_rate"], q["rate_class"])
    st = np.array([float(x[1]) for x in rows])
    assert np.allclose(st, op["stat"], rtol=2e-5, atol=2e-6) and [int(x[2]) for x in rows] == op["rcmin"].tolist()
    assert sum(int(x[6]) for x in rows) > 50000 * 100     # 100 000 null samples over 10 bins, most pairs scored
    pv = np.array([float(x[5]) if x[5] != "NA" else np.nan for x in rows])
    assert np.nanmin(pv) > 0 and np.nanmax(pv) <= 1
    out = run(tmp, "param=ProteinPairCompensation.bpp", "--seed=3", "statistic.output.file=comp.txt")
    _, rows = table(os.path.join(tmp, "comp.txt"))
    assert len(rows) == 129 * 128 // 2 and all(float(x[1]) <= 1 + 1e-12 for x in rows)
    for ex in ("ProteinGroupCorrelation", "ProteinGroupCompensation"):
        run(tmp, "param=" + ex + ".bpp", "--seed=3", "clustering.null.number=3")
        hdr, rows = table(os.path.join(tmp, "Myo_stats.csv"))
        assert hdr[:2] == ["Group", "Size"] and len(rows) > 20
        hdr, rows = table(os.path.join(tmp, "Myo_null.csv"))
        assert hdr[0] == "Rep" and {int(x[0]) for x in rows} == {0, 1, 2}
    out = run(tmp, "param=ProteinMappingOnly.bpp")
    _, rows = table(os.path.join(tmp, "Myo_counts.txt"))
    assert len(rows) == 197
    # its side outputs: marginal ancestral sequences (inner nodes by id, then the 100 observed sequences) and the tagged tree
    fa = open(os.path.join(tmp, "Myo_ancestors.fasta")).read().split(">")[1:]
    names = [x.split("\n", 1)[0] for x in fa]
    seqs = ["".join(x.split("\n")[1:]) for x in fa]
    p, o = dry_run(BIN, tmp, "param=ProteinMappingOnly.bpp")
    c = decode(o); c["parent"] = c["parent"].astype(np.int32)
    inner = [v for v in range(len(c["parent"])) if v in set(c["parent"].tolist())]
    assert names[:len(inner)] == [str(v) for v in inner] and len(names) == len(inner) + c["codes"].shape[0]
    assert all(len(x) == 129 for x in seqs)
    anc = O.ancestral_states(c["parent"], c["brlen"], c["Q"], c["pi"], c["rates"], c["probs"], c["codes"], c["code_mask"])
    aa = "ARNDCQEGHILKMFPSTWYV"
    exp = ["".join(aa[k] for k in anc[v]) for v in inner]
    assert sum(a != b for x, y in zip(seqs, exp) for a, b in zip(x, y)) <= 2
    tree = open(os.path.join(tmp, "Myo_tags.dnd")).read().strip()
    assert tree.endswith(")%d;" % (len(c["parent"]) - 1)) and tree.count("(") == len(inner)
    tl = open(os.path.join(tmp, "Myo_tags_tln.txt")).read().split("\n")
    assert tl[0] == "Name\tId" and len([x for x in tl[1:] if x]) == c["codes"].shape[0]


def test_mica_command_line(myo):
    """mica param=... (CoMap/Mica.cpp): the table with a model (norm-conditioned parametric bootstrap) equals the C ABI
    on the inputs the binary used; the z-score and nonparametric-bootstrap methods without a model run and fill their
    columns; APC / RCW are the products of the average MIs (Mica.cpp:661-662)."""
    from comap_b200 import api
    tmp, _ = myo
    mica = os.path.join(os.path.dirname(BIN), "mica_b200")
    common = ["alphabet=Protein", "input.sequence.file=Myoglobin.aln.sel.mase", "input.sequence.format=Mase",
              "input.sequence.sites_to_use=nogap", "input.remove_const=yes"]
    model = ["use_model=yes", "input.tree.file=Myo.dnd", "model=JTT92", "rate_distribution=Gamma(n=4, alpha=0.985435)"]
    p = subprocess.run([mica] + common + model + ["output.file=mica.txt", "null.method=parametric-bootstrap", "null.nb_rep_CPU=4",
                                                  "null.nb_rep_RAM=200", "null.nb_rate_classes=3", "null.output.file=mica.null.txt",
                                                  "--seed=9"], cwd=tmp, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-2000:]
    hdr, rows = table(os.path.join(tmp, "mica.txt"))
    assert hdr == ["Group", "MI", "APC", "RCW", "Hjoint", "Hmin", "Nmin", "Bs.p.value", "Bs.nb"] and len(rows) == 129 * 128 // 2
    c = host_inputs(tmp)
    ctx = api.Context(device=0)
    ctx.set_tree(c["parent"], c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
    ctx.set_alignment(c["codes"], c["code_mask"]); ctx.map()
    h, avg = ctx.mica_sites()
    ctx.mica_null_parametric(9, 4, 200, K=3)
    q = ctx.mica_pairs("nmin", use_null=True)
    co = c["coords"]
    assert [r[0] for r in rows] == ["[%d;%d]" % (co[i], co[j]) for i, j in zip(q["i"], q["j"])]
    assert [r[1] for r in rows] == [g(v) for v in q["mi"]] and [r[4] for r in rows] == [g(v) for v in q["hjoint"]]
    assert [r[5] for r in rows] == [g(v) for v in q["hmin"]] and [r[6] for r in rows] == [g(v) for v in q["nmin"]]
    assert [r[2] for r in rows] == [g(avg[i] * avg[j] / avg.mean()) for i, j in zip(q["i"], q["j"])]
    assert [r[3] for r in rows] == [g(avg[i] * avg[j] / 2.0) for i, j in zip(q["i"], q["j"])]
    assert [r[7] for r in rows] == ["NA" if np.isnan(v) else g(v) for v in q["pvalue"]] and [int(r[8]) for r in rows] == q["nsim"].tolist()
    nh, nrows = table(os.path.join(tmp, "mica.null.txt"))
    assert nh == ["MI", "Hjoint", "Hmin", "Nmin"] and len(nrows) == 800
    ctx.close()
    for method, extra in (("z-score", ["null.method_zscore.stat=MIp"]), ("nonparametric-bootstrap", ["null.nb_rep_CPU=5", "null.nb_rep_RAM=300"]),
                          ("none", [])):
        p = subprocess.run([mica] + common + ["output.file=m2.txt", "null.method=" + method, "null.nb_rate_classes=4", "--seed=3"] + extra,
                           cwd=tmp, capture_output=True, text=True)
        assert p.returncode == 0, p.stdout[-2000:]
        hdr, rows2 = table(os.path.join(tmp, "m2.txt"))
        want = ["Group", "MI", "APC", "RCW", "Hjoint", "Hmin"] + ([] if method == "none" else ["Bs.p.value", "Bs.nb"])
        assert hdr == want and len(rows2) == len(rows)
        assert [r[1] for r in rows2] == [r[1] for r in rows]            # the statistic does not depend on the model
        if method != "none":
            pv = np.array([float(r[6]) if r[6] != "NA" else np.nan for r in rows2])
            assert np.nanmin(pv) > 0 and np.nanmax(pv) <= 1 and (~np.isnan(pv)).mean() > 0.9
    p = subprocess.run([mica] + common + ["output.file=m3.txt", "null.method=parametric-bootstrap"], cwd=tmp, capture_output=True, text=True)
    assert p.returncode == 255 and "You need to specify a model" in p.stdout


def test_mica_examples_run_with_their_own_option_files(tmp_path):
    """examples/RNA/BacteriaSSU/options_{perm,npbs,pbs,zscore}.mica as shipped (README.md there): 760 complete variable
    sites x 40 taxa.  The permutation table is checked against the C ABI on the inputs the binary used."""
    from comap_b200 import api, build as b
    b.build_host()
    tmp = str(tmp_path)
    write_fixture(tmp, "bacteria_ssu")
    mica = os.path.join(os.path.dirname(BIN), "mica_b200")
    n = 760 * 759 // 2
    p = subprocess.run([mica, "param=options_perm.mica", "--seed=21"], cwd=tmp, capture_output=True, text=True)
    assert p.returncode == 0 and "Maximum number of permutations" in p.stdout, p.stdout[-2000:]
    hdr, rows = table(os.path.join(tmp, "Bacteria_SSU.MI_perm.sged"))
    assert hdr == ["Group", "MI", "APC", "RCW", "Hjoint", "Hmin", "Perm.p.value", "Perm.nb"] and len(rows) == n
    _, o = dry_run(BIN, tmp, "param=options_perm.mica")
    c = decode(o)
    # without a model mica keeps the alignment's sequence order (the dry run lists the rows in the tree's): the
    # statistics do not depend on it, the shuffles start from it
    row = {name: k for k, name in enumerate(o["leaves"])}
    c["codes"] = np.ascontiguousarray(c["codes"][[row[name] for name in o["sequences"]]])
    ctx = api.Context(device=0)
    ctx.set_tree(c["parent"].astype(np.int32), c["brlen"]); ctx.set_model(c["Q"], c["pi"], c["rates"], c["probs"])
    ctx.set_alignment(c["codes"], c["code_mask"])
    pv, nb = ctx.mica_permutations(21, 1000)
    q = ctx.mica_pairs("hmin")
    ctx.close()
    assert [r[1] for r in rows] == [g(v) for v in q["mi"]]
    assert [r[6] for r in rows] == [g(v) for v in pv] and [int(r[7]) for r in rows] == nb.tolist()
    assert nb.max() == 1000 and np.median(nb) < 200           # most pairs stop early, coevolving stems do not
    for opt, out, cols in (("options_npbs.mica", "Bacteria_SSU.MI_NPBS.sged", ["Bs.p.value", "Bs.nb"]),
                           ("options_zscore.mica", "Bacteria_SSU.MI_zscore.sged", ["Bs.p.value", "Bs.nb"]),
                           ("options_pbs.mica", "Bacteria_SSU.MI_PBS.sged", ["Nmin", "Bs.p.value", "Bs.nb"])):
        p = subprocess.run([mica, "param=" + opt, "--seed=21"], cwd=tmp, capture_output=True, text=True)
        assert p.returncode == 0, p.stdout[-2000:]
        hdr, rows2 = table(os.path.join(tmp, out))
        assert hdr == ["Group", "MI", "APC", "RCW", "Hjoint", "Hmin"] + cols and len(rows2) == n
        assert [r[1] for r in rows2] == [r[1] for r in rows]
        pcol = len(hdr) - 2
        pv2 = np.array([float(r[pcol]) if r[pcol] != "NA" else np.nan for r in rows2])
        assert np.nanmin(pv2) > 0 and np.nanmax(pv2) <= 1 and (~np.isnan(pv2)).mean() > 0.5


def test_error_exit_code_and_message(myo):
    tmp, _ = myo
    p = subprocess.run([BIN] + COMMON + ["analysis=pairwise", "statistic=Compensation"], cwd=tmp, capture_output=True, text=True)
    assert p.returncode == 255 and "Compensation distance must be used with a mapping procedure" in p.stdout
    p = subprocess.run([BIN] + COMMON + ["analysis=bogus"], cwd=tmp, capture_output=True, text=True)
    assert p.returncode == 255 and "Unknown analysis type" in p.stdout


def test_two_gpus_write_the_same_tables(myo):
    """comap_b200.gpus=2 (null replicates and pair rows sharded over two contexts of one process, NCCL
    all-gather inside the library; clustering-null replicates dealt to the GPUs): every output file is
    byte-identical to the one-GPU run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    tmp, golden = myo
    pair = ["analysis=pairwise", "statistic=Correlation", "statistic.null=yes", "statistic.null.nb_rep_CPU=5",
            "statistic.null.nb_rep_RAM=200", "statistic.null.nb_rate_classes=4"]
    run(tmp, *COMMON, *pair, "statistic.output.file=one.txt", "statistic.null.output.file=one_null.txt")
    run(tmp, *COMMON, *pair, "statistic.output.file=two.txt", "statistic.null.output.file=two_null.txt", "comap_b200.gpus=2")
    for a, b in (("one.txt", "two.txt"), ("one_null.txt", "two_null.txt")):
        assert open(os.path.join(tmp, a), "rb").read() == open(os.path.join(tmp, b), "rb").read(), a
    clu = ["analysis=clustering", "clustering.distance=cor", "clustering.method=complete", "clustering.null=yes",
           "clustering.null.number=3", "clustering.maximum_group_size=5"]
    run(tmp, *COMMON, *clu, "clustering.output.groups.file=g1.txt", "clustering.null.output.file=n1.txt")
    run(tmp, *COMMON, *clu, "clustering.output.groups.file=g2.txt", "clustering.null.output.file=n2.txt", "comap_b200.gpus=2")
    for a, b in (("g1.txt", "g2.txt"), ("n1.txt", "n2.txt")):
        assert open(os.path.join(tmp, a), "rb").read() == open(os.path.join(tmp, b), "rb").read(), a


def test_continuous_simulations_from_the_command_line(myo):
    """simulations.continuous = yes (CoMap.cpp:146) is accepted for Gamma rate distributions and changes the null."""
    tmp, golden = myo
    pair = ["analysis=pairwise", "statistic=Correlation", "statistic.null=yes", "statistic.null.nb_rep_CPU=3",
            "statistic.null.nb_rep_RAM=200", "statistic.null.nb_rate_classes=3"]
    out = run(tmp, *COMMON, *pair, "statistic.output.file=d.txt", "statistic.null.output.file=d_null.txt")
    assert "discrete" in out
    out = run(tmp, *COMMON, *pair, "statistic.output.file=c.txt", "statistic.null.output.file=c_null.txt",
              "simulations.continuous=yes")
    assert "continuous" in out
    hd, rd = table(os.path.join(tmp, "d_null.txt")); hc, rc = table(os.path.join(tmp, "c_null.txt"))
    assert hd == hc and len(rd) == len(rc) == 600 and rd != rc
    hs, rs = table(os.path.join(tmp, "c.txt"))
    assert hs[-2:] == ["PValue", "Nsim"] and len(rs) == 129 * 128 // 2
