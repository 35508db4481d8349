"""Summarises ncu outputs from gpurun_out/ into profiles/ (tracked).

    python tools/ncu_summary.py launches <launches.csv> <out.txt>
    python tools/ncu_summary.py kernel <report.ncu-rep> <out.txt>
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]); name = re.sub(r".*::", "", name)
        v = float(r[vi].replace(",", "")); u = r[ui]
        v = v / 1e3 if u == "us" else v / 1e6 if u == "ns" else v
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none`\n")
        f.write("# (cold-cache, serialised launches: compare SHARES, not absolutes); source: %s\n" % path)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-50s launches=%5d total_ms=%10.3f avg_ms=%9.4f share=%6.1f%%\n" % (k[:50], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
        f.write("total_ms=%.3f\n" % tot)


def kernel(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none; source: %s\n" % path)
        for vals in rows[2:]:
            d = dict(zip(hdr, zip(units, vals)))
            f.write("## %s\n" % d.get("Kernel Name", ("", "?"))[1][:120])
            for k in KEYS:
                if k in d:
                    f.write("%-70s %-12s %s\n" % (k, d[k][0], d[k][1]))
            for k in sorted(d):
                if "warp_issue_stalled" in k and k.endswith("_per_warp_active.pct"):
                    f.write("%-70s %-12s %s\n" % (k, d[k][0], d[k][1]))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
