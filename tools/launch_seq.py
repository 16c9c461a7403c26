"""Per-launch sequence of an ncu launch-list CSV (one timed step): id, kernel, grid, us, DRAM MB, tensor-pipe %."""
import csv, re, sys

def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    d = {}
    for x in csv.DictReader(lines):
        i = int(x["ID"])
        d.setdefault(i, {"k": x["Kernel Name"], "grid": x["Grid Size"]})[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
    seq = []
    for i in sorted(d):
        e = d[i]
        n = re.sub(r"\(.*", "", e["k"]).replace("void fdg::", "").replace("void ", "").replace("fdg::", "")
        seq.append((i, n[:58], e["grid"], e["gpu__time_duration.sum"] / 1e3,
                    (e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0)) / 1e6,
                    e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0)))
    return seq

if __name__ == "__main__":
    seq = load(sys.argv[1])
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    for s in seq:
        if s[3] >= thr:
            print("%4d %-58s %-14s %8.1f us %8.1f MB %5.1f" % s)
    print("total %.2f ms over %d launches" % (sum(s[3] for s in seq) / 1e3, len(seq)))
