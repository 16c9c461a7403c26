"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/make_summary.py launches gpurun_out/launches_X.csv profiles/rNN_launches.md "title"
    python profiles/make_summary.py kernel   gpurun_out/Y.ncu-rep      profiles/rNN_kernel_Y.md "title"
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst, title):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ik, iv, iu, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    rows = []
    for row in r:
        if len(row) <= iv:
            continue
        v = float(row[iv].replace(",", ""))
        v = v / 1e3 if row[iu] == "ns" else (v * 1e3 if row[iu] == "ms" else v)
        rows.append((row[ik].split("(")[0].replace("void ", ""), row[ig], v))
    tot = sum(x[2] for x in rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, g, v in rows:
        agg[k][0] += 1
        agg[k][1] += v
    out = ["# %s" % title, "",
           "Source: `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` around ONE timed step of",
           "`bench.py --steps 1 --warmup 3 --quick` (batch 16, 256x256, 1 GPU).  Per-launch times under ncu are cold-cache and",
           "serialised: compare SHARES, not absolutes.", "",
           "Total: %.1f ms over %d launches." % (tot / 1e3, len(rows)), "",
           "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        out.append("| `%s` | %d | %.2f | %.1f %% |" % (k[:90], v[0], v[1] / 1e3, 100 * v[1] / tot))
    fam = collections.defaultdict(float)
    for k, g, v in rows:
        key = ("tcgen05 conv (conv_umma / conv_halo)" if ("conv_umma" in k or "conv_halo" in k) else
               "tcgen05 wgrad (wgrad_umma / wgrad_halo)" if ("wgrad_umma" in k or "wgrad_halo" in k) else
               "fp32 SIMT conv / wgrad" if "simt" in k else
               "element-wise BN/ReLU backward" if "ew_bwd" in k else
               "torch (memset / foreach)" if ("at::" in k or "vectorized" in k or "elementwise" in k) else "other fdg kernels")
        fam[key] += v
    out += ["", "| family | ms | share |", "|---|---:|---:|"]
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]):
        out.append("| %s | %.2f | %.1f %% |" % (k, v / 1e3, 100 * v / tot))
    open(dst, "w").write("\n".join(out) + "\n")


KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]


def kernel(src, dst, title):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    md = ["# %s" % title, "", "Source: `ncu --set full --clock-control none --import-source on` (`%s`), read with `ncu -i ... --page raw --csv`." % src, ""]
    for row in r[2:]:
        name = row[hdr.index("Kernel Name")]
        md += ["## `%s`" % name[:120], "", "| metric | value |", "|---|---|"]
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                md.append("| %s | %s %s |" % (k, row[i], units[i]))
        md.append("")
    open(dst, "w").write("\n".join(md) + "\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else sys.argv[3])
