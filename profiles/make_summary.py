"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/make_summary.py launches gpurun_out/launches_X.csv profiles/rNN_launches.md "title"
    python profiles/make_summary.py kernel   gpurun_out/Y.ncu-rep      profiles/rNN_kernel_Y.md "title"
    python profiles/make_summary.py traffic  gpurun_out/launches_X.csv profiles/rNN_traffic.md "title" [profiles/rNN_traffic.json]
    python profiles/make_summary.py sass     fdgan_b200/libfdgan_b200.so profiles/rNN_sass_opcounts.md "title"
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



def traffic(src, dst_md, title, dst_json=None):
    """Launch list with per-launch DRAM bytes and tensor-pipe activity (ncu --metrics gpu__time_duration.sum,
    dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed)."""
    import json
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ik, im, iv, iu, iid = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    rows = collections.OrderedDict()
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for row in r:
        k = row[ik].split("(")[0].replace("void fdg::", "").replace("void ", "").replace("fdg::", "")
        d = rows.setdefault(row[iid], {"k": k})
        v = float(row[iv].replace(",", ""))
        u = row[iu]
        if row[im] == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        elif row[im].startswith("dram__bytes_read"):
            d["rd"] = v * scale[u]
        elif row[im].startswith("dram__bytes_write"):
            d["wr"] = v * scale[u]
        else:
            d["tc"] = v
    rows = list(rows.values())

    def famof(k):
        if "conv_umma" in k or "conv_halo" in k or "conv_k1" in k:
            return "conv_tcgen05"
        if "wgrad" in k:
            return "wgrad"
        if "ew_bwd" in k or "affine_accum" in k:
            return "ew_bwd"
        if "conv_simt" in k or "conv_thin" in k or "conv_cin1" in k:
            return "conv_simt_f32"
        if "freq" in k:
            return "freq"
        return "other"

    fam = collections.OrderedDict()
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
    for d in rows:
        a = fam.setdefault(famof(d["k"]), dict(launches=0, ms=0.0, dram=0.0, tc_ms=0.0))
        a["launches"] += 1
        a["ms"] += d["us"] / 1e3
        a["dram"] += d.get("rd", 0) + d.get("wr", 0)
        a["tc_ms"] += d.get("tc", 0) / 100 * d["us"] / 1e3
        b = agg[d["k"]]
        b[0] += 1; b[1] += d["us"]; b[2] += d.get("rd", 0); b[3] += d.get("wr", 0); b[4] += d.get("tc", 0) * d["us"]
    tot = sum(b[1] for b in agg.values())
    rd, wr = sum(d.get("rd", 0) for d in rows), sum(d.get("wr", 0) for d in rows)
    out = ["# %s" % title, "",
           "Source: `ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none` around ONE timed step of",
           "`bench.py --steps 1 --warmup 3 --quick` (batch 16, 256x256, 1 GPU).  Per-launch times under ncu are cold-cache and "
           "serialised: compare SHARES, not absolutes.", "",
           "Total: %.1f ms over %d launches; DRAM read %.1f GB + write %.1f GB per step (= %.1f ms at the measured 6547.5 GB/s)."
           % (tot / 1e3, len(rows), rd / 1e9, wr / 1e9, (rd + wr) / 6547.5e9 * 1e3), "",
           "| kernel | launches | total ms | share | DRAM read GB | DRAM write GB | DRAM TB/s | tensor pipe active % |",
           "|---|---:|---:|---:|---:|---:|---:|---:|"]
    for k, b in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        out.append("| `%s` | %d | %.2f | %.1f %% | %.2f | %.2f | %.2f | %.1f |"
                   % (k[:70], b[0], b[1] / 1e3, 100 * b[1] / tot, b[2] / 1e9, b[3] / 1e9, (b[2] + b[3]) / b[1] / 1e6, b[4] / b[1]))
    out += ["", "| family | launches | ms | DRAM GB | mean DRAM MB / launch | tensor pipe active % |", "|---|---:|---:|---:|---:|---:|"]
    for k, v in fam.items():
        out.append("| %s | %d | %.2f | %.1f | %.1f | %.1f |" % (k, v["launches"], v["ms"], v["dram"] / 1e9, v["dram"] / v["launches"] / 1e6,
                                                               100 * v["tc_ms"] / v["ms"]))
    with open(dst_md, "w") as f:
        f.write("\n".join(out) + "\n")
    if dst_json:
        js = {"src_sha": kernel_source_sha(),
              "source": "ncu launch list of one step of bench.py --steps 1 --warmup 3 --quick (batch 16, 256x256, 1 B200): " + src,
              "families": {k: dict(launches=v["launches"], ms_under_ncu=round(v["ms"], 3), dram_bytes_per_launch=round(v["dram"] / v["launches"]),
                                   dram_gbs_under_ncu=round(v["dram"] / v["ms"] / 1e6, 1), tensor_pipe_active_pct=round(100 * v["tc_ms"] / v["ms"], 1))
                           for k, v in fam.items()}}
        with open(dst_json, "w") as f:
            json.dump(js, f, indent=1)


def kernel_source_sha():
    """Same hash as bench.py:kernel_source_sha (ties a capture to the kernel sources it was taken on)."""
    import glob
    import hashlib
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(root, "fdgan_b200", "csrc", "*.cu*")) + glob.glob(os.path.join(root, "include", "*.h"))):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


SASS_OPS = ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "LDGSTS", "HMMA", "FFMA", "RED", "ATOM")


def sass(lib, dst, title):
    """Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md), from `cuobjdump -sass`."""
    import re
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
    names = iter(filt)
    rows = []
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = {"name": next(names), "n": 0}
            cur.update({k: 0 for k in SASS_OPS})
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            cur["n"] += 1
            op = m.group(1).split(".")[0]
            if op in cur:
                cur[op] += 1
    md = ["# %s" % title, "", "Source: `cuobjdump -sass %s` (kernel sources %s), counted by `profiles/make_summary.py sass`." % (lib, kernel_source_sha()),
          "`UTCHMMA` = tcgen05.mma (kind::f16), `LDTM` = tcgen05.ld (TMEM load), `UTMALDG` / `UTMASTG` = TMA tensor load / store,",
          "`UBLKCP` = cp.async.bulk (1-D bulk copy), `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, `LDGSTS` = cp.async.", "",
          "| kernel | SASS instr | " + " | ".join(SASS_OPS) + " |", "|---|---:|" + "---:|" * len(SASS_OPS)]
    for r in sorted(rows, key=lambda r: (-r["UTCHMMA"], -r["n"])):
        nm = re.sub(r"^void ", "", r["name"]).split("(")[0]
        md.append("| `%s` | %d | %s |" % (nm[:80], r["n"], " | ".join(str(r[k]) for k in SASS_OPS)))
    open(dst, "w").write("\n".join(md) + "\n")


if __name__ == "__main__":
    (traffic(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else None) if sys.argv[1] == "traffic" else
     {"launches": launches, "kernel": kernel, "sass": sass}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else sys.argv[3]))
