#!/usr/bin/env python
"""bench.py -- CoMap hot path on B200: site-pairs scored per second, mapping + null included.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], SURVEY.md s8d "config 4"): synthetic 5,000-site x
500-taxon nucleotide alignment, HKY85(kappa=2.5, pi=.3/.2/.2/.3) + Gamma(4, alpha=0.5),
random-join tree with Exp(0.02) branches (numpy seed 20251018); one step =
  map the alignment -> parametric-bootstrap null (1000 x 1000 paired simulated sites:
  2,000,000 sites simulated + mapped, 1,000,000 null statistics, binned by Nmin and sorted)
  -> all 12,497,500 site pairs scored (correlation) with p-values.
pairs per step = S(S-1)/2 + rep_cpu*rep_ram.

`value`  : device-resident throughput (inputs in HBM, results left in HBM), CUDA events.
`e2e`    : same step through the C ABI with host buffers: alignment H2D from pinned memory,
           per-site columns and all 8 pair-table columns D2H into pinned memory.
`roofline`: the dominant kernel family (K1 mapping pass: down + finish + up launches over
           one batch of sites), algorithmic bytes per site from SURVEY.md s8(d).
`cpu_baseline` / --impl reference: the CPU oracle port of the reference's algorithm
           (the upstream binary needs Bio++ and cannot be built here) on a bounded sample,
           extrapolated linearly to the full step and labelled as such.  The reference is
           single-threaded; like its users we run one process per host core (shards of the
           null replicates and pair rows), all at once, and add their throughputs.
Multi-GPU: strong scaling of the same job -- null replicates and pair rows are sharded,
null samples are all-gathered with NCCL, every rank bins/sorts the union.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from comap_b200 import synthetic as syn  # noqa: E402

CFG = dict(sites=5000, taxa=500, kappa=2.5, pi=[0.3, 0.2, 0.2, 0.3], alpha=0.5, classes=4,
           mean_brlen=0.02, tree_seed=20251018, aln_seed=1, null_seed=2, rep_cpu=1000, rep_ram=1000,
           null_bins=10, statistic="correlation")


def workload(cfg):
    parent, brlen = syn.random_tree(cfg["taxa"], cfg["tree_seed"], cfg["mean_brlen"])
    Q, pi = syn.hky85(cfg["kappa"], cfg["pi"])
    rates, probs = syn.gamma_rates(cfg["alpha"], cfg["classes"])
    return dict(parent=parent, brlen=brlen, Q=Q, pi=pi, rates=rates, probs=probs,
                code_mask=syn.identity_code_mask(4))


def algorithmic_bytes_per_site(T, C, A, B):
    """tips + inner down-partials written once and read once + output."""
    return T + 2 * (T - 3) * C * A * 8 + 8 * B + 28


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.rows.append(ln.strip().split(", "))

    def stop(self):
        if not self.p:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[4 + k].strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample
# ------------------------------------------------------------------------------------------
def cpu_sample(cfg, w, aln_codes=None, sample_sites=800, sample_ram=3000, seed=0):
    """Times the oracle (1 thread) on a sample of the step and extrapolates linearly.

    Components timed: (a) simulate+map+paired statistic for 1 outer replicate of
    `sample_ram` site pairs, (b) mapping of `sample_sites` observed sites, (c) all pairs of
    those sites with the p-value scan against a null of the full step's size.
    """
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as O
    S, R, RC, K = cfg["sites"], cfg["rep_ram"], cfg["rep_cpu"], cfg["null_bins"]
    t0 = time.perf_counter()
    s1, _ = O.simulate(w["parent"], w["brlen"], w["Q"], w["pi"], w["rates"], w["probs"], cfg["null_seed"] + seed, 0, sample_ram)
    s2, _ = O.simulate(w["parent"], w["brlen"], w["Q"], w["pi"], w["rates"], w["probs"], cfg["null_seed"] + seed, sample_ram, sample_ram)
    t_sim = time.perf_counter() - t0
    t0 = time.perf_counter()
    nl = O.null_intra(w["parent"], w["brlen"], w["Q"], w["pi"], w["rates"], w["probs"], cfg["statistic"],
                      s1[None], s2[None], K, 10.0)
    t_null = time.perf_counter() - t0
    if aln_codes is None:
        aln_codes, _ = O.simulate(w["parent"], w["brlen"], w["Q"], w["pi"], w["rates"], w["probs"], cfg["aln_seed"], 0, sample_sites)
    t0 = time.perf_counter()
    m = O.map_sites(w["parent"], w["brlen"], w["Q"], w["pi"], w["rates"], w["probs"], aln_codes[:, :sample_sites], w["code_mask"])
    t_map = time.perf_counter() - t0
    # null of the full step's size for the p-value scan: resample the small null
    rng = np.random.default_rng(seed)
    raw = nl["raw"]
    ok = ~np.isnan(raw[:, 0])
    nmax = float(m["norm"].max())
    big = rng.choice(np.flatnonzero(ok), size=RC * R, replace=True)
    cat = np.minimum((raw[big, 3] / (nmax / K)).astype(np.int64), K)
    order = np.lexsort((raw[big, 0], cat))
    keep = cat[order] < K
    srt = raw[big, 0][order][keep]
    offs = np.searchsorted(cat[order][keep], np.arange(K + 1))
    t0 = time.perf_counter()
    pr = O.pairs(cfg["statistic"], m["n"], m["norm"], m["post_rate"], m["rate_class"], null=(K, nmax, offs, srt))
    t_pairs = time.perf_counter() - t0
    n_pairs_s = len(pr["i"])
    per_null_site = (t_sim + t_null) / (2 * sample_ram)   # simulate + map (+ paired stat) per simulated site
    per_obs_site = t_map / sample_sites
    per_obs_pair = t_pairs / max(1, n_pairs_s)
    full = per_null_site * 2 * RC * R + per_obs_site * S + per_obs_pair * (S * (S - 1) // 2)
    total_pairs = S * (S - 1) // 2 + RC * R
    sample_pairs = n_pairs_s + sample_ram
    sample_time = t_sim + t_null + t_map + t_pairs
    return dict(value=total_pairs / full, full_step_seconds=full, sample_seconds=sample_time,
                sample_pairs_per_s=sample_pairs / sample_time,
                sample=("oracle port, 1 thread: simulate+map+pair %d null site pairs (%.2fs), map %d observed sites "
                        "(%.2fs), score their %d pairs with p-value scan against a %d-sample null (%.2fs); "
                        "extrapolated linearly to %d null pairs + %d observed sites + %d pairs"
                        % (sample_ram, t_sim + t_null, sample_sites, t_map, n_pairs_s, RC * R, t_pairs, RC * R, S,
                           S * (S - 1) // 2)))


def _cpu_worker(job):
    cfg, w, aln_codes, seed = job
    return cpu_sample(cfg, w, aln_codes=aln_codes, seed=seed)


def cpu_sample_all_cores(cfg, w, aln_codes=None, seed=0, n_proc=None):
    """The only parallelism the reference admits is independent processes (shards of the outer null
    replicates and of the pair rows): one oracle sample per host core, all running at the same time;
    the job's throughput is the sum of the per-process throughputs measured under that load."""
    import multiprocessing as mp
    n_proc = n_proc or os.cpu_count() or 1
    if n_proc == 1:
        r = cpu_sample(cfg, w, aln_codes=aln_codes, seed=seed)
        return dict(r, cores=1)
    ctx = mp.get_context("spawn")  # the parent may hold a CUDA context
    with ctx.Pool(n_proc) as pool:
        rs = pool.map(_cpu_worker, [(cfg, w, aln_codes, seed * 1000 + i) for i in range(n_proc)])
    value = float(sum(r["value"] for r in rs))
    total_pairs = cfg["sites"] * (cfg["sites"] - 1) // 2 + cfg["rep_cpu"] * cfg["rep_ram"]
    return dict(value=value, full_step_seconds=total_pairs / value, cores=n_proc,
                sample_seconds=float(max(r["sample_seconds"] for r in rs)),
                sample_pairs_per_s=float(sum(r["sample_pairs_per_s"] for r in rs)),
                sample="%d concurrent processes, each: %s" % (n_proc, rs[0]["sample"]))


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(cfg)
    for i in range(args.warmup):
        cpu_sample(cfg, w, sample_sites=16, sample_ram=16, seed=100 + i)
    vals, secs, last = [], [], None
    for i in range(args.steps):
        last = cpu_sample_all_cores(cfg, w, seed=i)
        vals.append(last["value"]); secs.append(last["full_step_seconds"])
    v = float(np.mean(vals))
    line = dict(impl="reference", metric="site_pairs_scored_per_s_incl_mapping_and_null", value=v, unit="pairs/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=float(np.mean(secs) * 1e3),
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                config=config_dict(cfg, args.gpus),
                cpu_baseline=dict(value=v, unit="pairs/s", cores=last["cores"], kind="port", sample=last["sample"],
                                  host_cores=os.cpu_count(), sample_pairs_per_s=last["sample_pairs_per_s"],
                                  note="upstream CoMap needs Bio++ >= 3.0 (not installable offline); this is the "
                                       "CPU oracle restatement, extrapolated from the sample"),
                e2e=dict(value=v, unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def config_dict(cfg, n_gpus):
    return dict(workload="configs[3]: synthetic %d-site x %d-taxon nucleotide alignment, HKY85+Gamma4, all-pairs "
                         "%s + %dx%d null" % (cfg["sites"], cfg["taxa"], cfg["statistic"], cfg["rep_cpu"], cfg["rep_ram"]),
                sites=cfg["sites"], taxa=cfg["taxa"], branches=2 * cfg["taxa"] - 3, rate_classes=cfg["classes"],
                rep_cpu=cfg["rep_cpu"], rep_ram=cfg["rep_ram"], null_bins=cfg["null_bins"],
                pairs_per_step=cfg["sites"] * (cfg["sites"] - 1) // 2 + cfg["rep_cpu"] * cfg["rep_ram"],
                parallelism="null replicates + pair rows sharded over %d GPU(s); NCCL all-gather of null samples" % n_gpus,
                l2="per-step working set (down-partials of each null batch, GBs) exceeds the 126 MB L2; no explicit flush")


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class _DevArr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = dict(shape=(n,), typestr="<f8", data=(ptr, False), version=2)


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from comap_b200 import api, parallel as par

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; comap_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    w = workload(cfg)
    S, T, R, RC, K = cfg["sites"], cfg["taxa"], cfg["rep_ram"], cfg["rep_cpu"], cfg["null_bins"]
    B = 2 * T - 3
    stat = cfg["statistic"]
    # shard of the null replicates / pair rows owned by this rank
    bounds = par.replicate_bounds(RC, world)
    r0, r1 = bounds[rank]
    max_reps = max(e - b for b, e in bounds)

    with torch.cuda.stream(stream):
        ctx = api.Context(device=local, stream=stream.cuda_stream)
        ctx.set_tree(w["parent"], w["brlen"])
        ctx.set_model(w["Q"], w["pi"], w["rates"], w["probs"])
        codes, _ = ctx.simulate(cfg["aln_seed"], 0, S)     # the synthetic alignment (project's own simulator)
        pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
        codes_pin = pin((T, S), torch.uint8); codes_pin[:] = codes
        ctx.set_alignment(codes_pin, w["code_mask"])
        n_own = par.owned_pairs(S, rank, world)
        cols_pin = [pin((max(1, n_own),), {np.int32: torch.int32, np.float64: torch.float64, np.int64: torch.int64}[dt])
                    for dt in api.Context.COL_DTYPE]

        def null_dist():
            if world == 1:
                ctx.null_intra(stat, cfg["null_seed"], RC, R, K=K, nmax=-1.0)
            else:
                ctx.null_intra(stat, cfg["null_seed"], RC, R, K=0, rep_begin=r0, rep_end=r1)
                sp, mp, n = ctx.null_samples_dev()
                st_all, nm_all = par.all_gather_null(torch.as_tensor(_DevArr(sp, max(n, 1)), device="cuda"),
                                                     torch.as_tensor(_DevArr(mp, max(n, 1)), device="cuda"), n, max_reps * R)
                stream.synchronize()
                ctx.null_load_dev(st_all.data_ptr(), nm_all.data_ptr(), st_all.numel(), K, -1.0)

        def step_resident():
            ctx.map(want_vectors=False)
            null_dist()
            return ctx.pairs_resident(stat, use_null=True, shard_index=rank, shard_count=world)

        def step_e2e():
            # the six null-independent columns are scored first and copied out while the null
            # distribution is simulated; PValue / Nsim follow once the null exists
            ctx.set_alignment(codes_pin, w["code_mask"])
            ctx.map(want_vectors=False)
            ctx.pairs_resident(stat, use_null=False, shard_index=rank, shard_count=world, columns=0x3F)
            for k in range(6):
                ctx.pairs_fetch(k, cols_pin[k])
            null_dist()
            n = ctx.pairs_resident(stat, use_null=True, shard_index=rank, shard_count=world, columns=0xC0)
            for k in (6, 7):
                ctx.pairs_fetch(k, cols_pin[k])
            ctx.sync()
            return n

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(fn, steps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())

        for _ in range(max(3, args.warmup)):
            n_rows = step_resident()
        step_e2e()

        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ctx.profile_reset(); ctx.profile_enable(True)
        l0 = ctx.launch_count()
        ms = timed(step_resident, args.steps)
        launches = ctx.launch_count() - l0
        prof = {k: ctx.profile_get(k) for k in ("map_down", "map_up", "simulate", "null_pairs", "sort", "pairs")}
        ctx.profile_enable(False)
        ms_e2e = timed(step_e2e, args.steps)
        clocks = sampler.stop() if rank == 0 else None

        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(lt)
        pairs_per_step = S * (S - 1) // 2 + RC * R
        if rank == 0:
            # roofline of the dominant kernel family: K1 mapping passes
            map_ms = prof["map_down"][0] + prof["map_up"][0]
            sites_mapped = args.steps * (S + 2 * (r1 - r0) * R)
            abytes_site = algorithmic_bytes_per_site(T, cfg["classes"], 4, B)
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except Exception:
                pass
            peak = float(peaks.get("hbm_gbs", 6650.0))
            achieved = sites_mapped * abytes_site / (map_ms * 1e-3) / 1e9 if map_ms > 0 else 0.0
            n_pass = max(1, prof["map_up"][1])
            # DRAM bytes actually moved, from the committed `ncu --set full` captures of one pass over
            # 128 256 sites (profiles/r1j_k1_{down,up}_mma.txt: dram__bytes_read.sum + dram__bytes_write.sum);
            # below the algorithmic figure because cherry partials are recomputed, not stored
            traffic_site = (69.0e6 + 5.2964e9 + 5.6296e9 + 1.4807e9) / 128256.0
            line = dict(metric="site_pairs_scored_per_s_incl_mapping_and_null",
                        value=pairs_per_step * args.steps / (ms * 1e-3), unit="pairs/s", n_gpus=world, steps=args.steps,
                        warmup=max(3, args.warmup), ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong",
                        vs_baseline=None, dtype="f64", data="synthetic", config=config_dict(cfg, world), clocks=clocks,
                        e2e=dict(value=pairs_per_step * args.steps / (ms_e2e * 1e-3), unit="pairs/s",
                                 ms_per_step=ms_e2e / args.steps,
                                 h2d_bytes_per_step=int(T * S + 4 * 256),
                                 d2h_bytes_per_step=int(sum(c.itemsize for c in cols_pin) * n_rows + S * 8 * 4)),
                        gpu_launches=int(lt.item()),
                        roofline=dict(bound="hbm", kernel="K1 mapping pass (k1_down + k1_finish + k1_up over one batch of sites)",
                                      achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                                      peak_source="MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                                      algorithmic_bytes_per_site=abytes_site, sites_per_step=sites_mapped // args.steps,
                                      avg_pass_ms=map_ms / n_pass, passes=int(n_pass),
                                      traffic=traffic_site * sites_mapped / n_pass, traffic_unit="bytes per pass",
                                      traffic_source="ncu dram bytes of k1_down_mma + k1_up_mma, profiles/r1j_*.txt, scaled by sites"),
                        kernel_ms_per_step={k: v[0] / args.steps for k, v in prof.items()})
            if world == 1 and not args.no_cpu_baseline:
                cb = cpu_sample_all_cores(cfg, w, aln_codes=codes)
                line["cpu_baseline"] = dict(value=cb["value"], unit="pairs/s", cores=cb["cores"], kind="port", sample=cb["sample"],
                                            host_cores=os.cpu_count(), sample_pairs_per_s=cb["sample_pairs_per_s"])
            print(json.dumps(line), flush=True)
        ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    for k in ("sites", "taxa", "rep_cpu", "rep_ram"):
        ap.add_argument("--" + k.replace("_", "-"), type=int, default=None)
    args = ap.parse_args()
    cfg = dict(CFG)
    for k in ("sites", "taxa", "rep_cpu", "rep_ram"):
        if getattr(args, k) is not None:
            cfg[k] = getattr(args, k)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
